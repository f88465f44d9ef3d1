// march_irk4_tm.cuh — the IRK4 contour march (march_irk4.cuh: the reference's 2-stage Gauss-Legendre stepper as ONE complex
// symmetric tridiagonal solve per step; scft.cc:671-693, drivescft.cc:130-146) for the benchmarked shape — uniform mesh,
// 513..1024 unknowns — with its loop-invariant complex coefficients in TENSOR MEMORY.
//
// Same algorithm and arithmetic as march_irk4_kernel<4,256,true> (218 registers, one CTA of 8 warps per SM, 34 % of the fp64
// pipe, latency-bound: 1.52e11 DOF-steps/s on the 4096-problem sweep).  Two tensor-memory versions were built:
//   C = 4, T = 256, 56 doubles per thread in TMEM, 128 registers, 2 CTAs = 16 warps per SM: 1.89e11.  With 16 warps that kernel is
//     ISSUE-bound (406 SASS instructions per warp-step of 128 nodes, 74 of them 32-bit shuffles of the complex cyclic
//     reduction and butterfly; ncu: issue slots 49 %, fp64 pipe 42 %) — occupancy was not the lever, instructions per node are.
//   C = 8, T = 128 (this file): one separator per 8 nodes halves the level-2/3 share; 97 doubles per thread in TMEM (196 of the
//     256 allocated columns), up to 255 registers for the state and the blocks in flight, 2 CTAs = 8 warps per SM: **2.68e11**
//     (+76 %, 33 % of the HBM roofline; the fp64 pipe caps this scheme near 50 %).
// The eight per-warp level-3 constants come from a conflict-free shared-memory table; the step is instantiated per quadrature
// phase (store / middle / pairing).  A throughput kernel: the tcgen05.ld waits lengthen a lone CTA's step, so engines with
// max_batch <= 148 keep the register-resident kernel (engine.cu).
#pragma once
#include "march1d_tmem.cuh"
#include "march_irk4.cuh"

namespace scftb {

constexpr int TM4_COLS = 256;

__device__ __forceinline__ void tm4_alloc(uint32_t *smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(TM4_COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm4_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TM4_COLS) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <int NR>
__device__ __forceinline__ cx tm_getc(const uint32_t (&r)[NR], int i) { return mk(tm_get(r, 2 * i), tm_get(r, 2 * i + 1)); }
template <int NR>
__device__ __forceinline__ void tm_putc(uint32_t (&r)[NR], int i, cx v) { tm_put(r, 2 * i, v.re); tm_put(r, 2 * i + 1, v.im); }

constexpr int C3S = 10;  // doubles per row of the level-3 constant table: 80-byte rows are 16-byte aligned and the eight rows a warp reads fall into disjoint banks

// TMEM column map of a thread (a double is two 32-bit columns, a complex number four):
//   block A  [  0, 80)  al[1..6], ca[0..6], be[0..5], sl                        (complex)
//   block B  [ 80,128)  su, A_off, -, pa[0..4], pg[0..4]                        (complex but A_off)
//   block C  [128,196)  binv, GL, GR, gl[0..6], gr[0..6]                        (complex)
constexpr int TM8_A = 0, TM8_B = 80, TM8_C = 128;
#ifndef IRK4_TWOSIDED
// 1: two-sided ("burn at both ends") elimination of the 7-node chunk: nodes 0..2 are eliminated downwards, nodes 6..4 upwards,
// both meet at node 3, substitution runs outwards from there.  Same operation and coefficient count as the one-sided sweeps,
// but the dependent chain per step is 2 + 1 + 1 + 3 complex operations instead of 6 + 1 + 6.  Measured (same parity tests green):
// 2.617e11 vs 2.680e11 DOF-steps/s for the one-sided sweeps — the chunk chain is not what the step waits on — so it stays off.
#define IRK4_TWOSIDED 0
#endif
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// store / load NC complex numbers (4 columns each) at taddr, in pieces of 16 and 4 columns
template <int NC>
__device__ __forceinline__ void tm_store_cx(uint32_t taddr, const cx (&v)[NC]) {
  static_assert(NC % 1 == 0, "");
  int done = 0;
#pragma unroll
  for (; done + 4 <= NC; done += 4) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 4; i++) tm_putc(w, i, v[done + i]);
    tm_st16(taddr + 4 * done, w);
  }
#pragma unroll
  for (; done < NC; done++) {
    uint32_t w[4];
    tm_put(w, 0, v[done].re); tm_put(w, 1, v[done].im);
    tm_st4(taddr + 4 * done, w);
  }
}

__global__ void __launch_bounds__(128, 2) march_irk4_tm_kernel(MarchParams P) {
  constexpr int C = 8, T = 128, CI = 7, NW = 4, SL = T * C;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  __shared__ cx s_ex[2][T];
  __shared__ cx s_l3[NW][9];       // P, D, Nx, GL0, GR0, GL30, GR30, cAu(real), csu
  __shared__ __align__(16) cx s_minv[NW][NW];
  __shared__ __align__(16) double s_c3[NW][C3S];   // per separator v: P.re, P.im, su.re, su.im, Nx.re, Nx.im, A_up
  __shared__ __align__(16) double s_pub[2][NW][PUBC];
  __shared__ double s_red[NW];
  __shared__ uint32_t s_tm;
  const int n = P.nsteps;
  const double dt = 1.0 / n;
  const cx z1 = mk(3.0, 1.7320508075688772);
  const double a_re = 12.0, a_im = 12.0 * 1.7320508075688772;

  if (wid == 0) tm4_alloc(&s_tm);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t tb = s_tm + ((uint32_t)(wid * 32) << 16);

  for (int p = blockIdx.x; p < P.nprob; p += gridDim.x) {
    if (P.skip && P.skip[p]) continue;
    const int pp = P.pshare ? 0 : p;
    const double L = P.L[pp];
    {
      auto wrow = [&](const Row &r, cx &wl, cx &wd, cx &wu) {   // W = z1 A + dt D
        wl = mk(fma(dt, r.Dl, z1.re * r.Al), z1.im * r.Al);
        wd = mk(fma(dt, r.Dd, z1.re * r.Ad), z1.im * r.Ad);
        wu = mk(fma(dt, r.Du, z1.re * r.Au), z1.im * r.Au);
      };
      cx ca[CI], al[CI], be[CI], gl[CI], gr[CI];
      double sAd;
      cx sl, sd, su;
      {
        Row rs = assemble_row(P, p, t * C + CI, L, dt);
        sAd = (rs.Al != 0.0) ? rs.Al : rs.Au;
        if (t * C + CI >= P.ni) { sl = mk(0.0); sd = mk(1.0); su = mk(0.0); }
        else wrow(rs, sl, sd, su);
      }
#if IRK4_TWOSIDED
      {
        //   al[1..6] <- the six elimination multipliers m1, m2, m4, m5, mL3, mR3
        //   ca[k]    <- A_off / pivot_k
        //   be[0..5] <- substitution couplings d0, d1, d2 (to node k+1), d4, d5, d6 (to node k-1)
        cx wl[CI], wd[CI], wu[CI], piv[CI], mL[CI], mR[CI];
        double aoff[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) {
          Row r = assemble_row(P, p, t * C + k, L, dt);
          if (t * C + k >= P.ni) { wl[k] = mk(0.0); wd[k] = mk(1.0); wu[k] = mk(0.0); }
          else wrow(r, wl[k], wd[k], wu[k]);
          aoff[k] = (r.Al == 0.0) ? r.Au : r.Al;
          mL[k] = mk(0.0); mR[k] = mk(0.0);
        }
        piv[0] = wd[0];
#pragma unroll
        for (int k = 1; k <= 2; k++) { mL[k] = wl[k] * cinv(piv[k - 1]); piv[k] = nfma(mL[k], wu[k - 1], wd[k]); }
        piv[6] = wd[6];
#pragma unroll
        for (int k = 5; k >= 4; k--) { mR[k] = wu[k] * cinv(piv[k + 1]); piv[k] = nfma(mR[k], wl[k + 1], wd[k]); }
        mL[3] = wl[3] * cinv(piv[2]); mR[3] = wu[3] * cinv(piv[4]);
        piv[3] = nfma(mR[3], wl[4], nfma(mL[3], wu[2], wd[3]));
        cx pinv[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) pinv[k] = cinv(piv[k]);
        auto solve = [&](const cx (&rhs)[CI], cx (&x)[CI]) {   // exact chunk solve with these factors
          cx y[CI];
          y[0] = rhs[0]; y[1] = nfma(mL[1], y[0], rhs[1]); y[2] = nfma(mL[2], y[1], rhs[2]);
          y[6] = rhs[6]; y[5] = nfma(mR[5], y[6], rhs[5]); y[4] = nfma(mR[4], y[5], rhs[4]);
          y[3] = nfma(mR[3], y[4], nfma(mL[3], y[2], rhs[3]));
          x[3] = y[3] * pinv[3];
          x[2] = nfma(wu[2], x[3], y[2]) * pinv[2]; x[1] = nfma(wu[1], x[2], y[1]) * pinv[1]; x[0] = nfma(wu[0], x[1], y[0]) * pinv[0];
          x[4] = nfma(wl[4], x[3], y[4]) * pinv[4]; x[5] = nfma(wl[5], x[4], y[5]) * pinv[5]; x[6] = nfma(wl[6], x[5], y[6]) * pinv[6];
        };
        {   // spikes: W_loc gl = Wl_first e_first, W_loc gr = Wu_last e_last
          cx e[CI] = {wl[0], mk(0.0), mk(0.0), mk(0.0), mk(0.0), mk(0.0), mk(0.0)};
          solve(e, gl);
          cx f[CI] = {mk(0.0), mk(0.0), mk(0.0), mk(0.0), mk(0.0), mk(0.0), wu[6]};
          solve(f, gr);
        }
        al[0] = mk(0.0); al[1] = mL[1]; al[2] = mL[2]; al[3] = mR[4]; al[4] = mR[5]; al[5] = mL[3]; al[6] = mR[3];
#pragma unroll
        for (int k = 0; k < CI; k++) ca[k] = pinv[k] * aoff[k];
        be[0] = wu[0] * pinv[0]; be[1] = wu[1] * pinv[1]; be[2] = wu[2] * pinv[2];
        be[3] = wl[4] * pinv[4]; be[4] = wl[5] * pinv[5]; be[5] = wl[6] * pinv[6];
        be[6] = mk(0.0);
      }
#else
      {
        cx Tl0 = mk(0.0), TuL = mk(0.0), pinv_prev = mk(0.0), Wu_prev = mk(0.0);
        cx alo[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) {
          Row r = assemble_row(P, p, t * C + k, L, dt);
          cx wl, wd, wu;
          if (t * C + k >= P.ni) { wl = mk(0.0); wd = mk(1.0); wu = mk(0.0); }
          else wrow(r, wl, wd, wu);
          cx piv = (k == 0) ? wd : wd - (wl * pinv_prev) * Wu_prev;
          cx pinv = cinv(piv);
          ca[k] = pinv * ((r.Al == 0.0) ? r.Au : r.Al);
          alo[k] = (k == 0) ? mk(0.0) : pinv * wl;
          al[k] = (k == 0) ? mk(0.0) : wl * pinv_prev;
          be[k] = (k == CI - 1) ? mk(0.0) : pinv * wu;
          if (k == 0) Tl0 = pinv * wl;
          if (k == CI - 1) TuL = pinv * wu;
          pinv_prev = pinv; Wu_prev = wu;
        }
        cx y[CI];
        y[0] = Tl0;
#pragma unroll
        for (int k = 1; k < CI; k++) y[k] = -(alo[k] * y[k - 1]);
        gl[CI - 1] = y[CI - 1];
#pragma unroll
        for (int k = CI - 2; k >= 0; k--) gl[k] = nfma(be[k], gl[k + 1], y[k]);
        gr[CI - 1] = TuL;
#pragma unroll
        for (int k = CI - 2; k >= 0; k--) gr[k] = -(be[k] * gr[k + 1]);
      }
#endif
      {   // block A: al[1..6], ca[0..6], be[0..5], sl
        cx blk[20];
#pragma unroll
        for (int k = 0; k < 6; k++) blk[k] = al[k + 1];
#pragma unroll
        for (int k = 0; k < 7; k++) blk[6 + k] = ca[k];
#pragma unroll
        for (int k = 0; k < 6; k++) blk[13 + k] = be[k];
        blk[19] = sl;
        tm_store_cx(tb + TM8_A, blk);
      }
      cx a, b, c;
      __syncthreads();
      s_ex[0][t] = gl[0]; s_ex[1][t] = gr[0];
      __syncthreads();
      {
        cx gl0n = (t + 1 < T) ? s_ex[0][t + 1] : mk(0.0), gr0n = (t + 1 < T) ? s_ex[1][t + 1] : mk(0.0);
        a = -(sl * gl[CI - 1]);
        b = nfma(su, gl0n, nfma(sl, gr[CI - 1], sd));
        c = -(su * gr0n);
      }
      const cx l3P = a, l3D = b, l3N = c;
      const cx A0 = (lane == 0) ? a : mk(0.0), C30 = (lane == 30) ? c : mk(0.0);
      if (lane == 31) { a = mk(0.0); b = mk(1.0); c = mk(0.0); }
      if (lane == 0) a = mk(0.0);
      if (lane == 30) c = mk(0.0);
      cx pa_[5], pg_[5];
#pragma unroll
      for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        cx am = shfl_up_c(a, d), bm = shfl_up_c(b, d), cm = shfl_up_c(c, d);
        cx ap = shfl_dn_c(a, d), bp = shfl_dn_c(b, d), cp = shfl_dn_c(c, d);
        cx alpha = (lane >= d) ? -(a * cinv(bm)) : mk(0.0);
        cx gamma = (lane + d <= 31) ? -(c * cinv(bp)) : mk(0.0);
        if (lane < d) { am = mk(0.0); cm = mk(0.0); }
        if (lane + d > 31) { ap = mk(0.0); cp = mk(0.0); }
        b = pfma(gamma, ap, pfma(alpha, cm, b));
        a = alpha * am;
        c = gamma * cp;
        pa_[s] = alpha; pg_[s] = gamma;
      }
      const cx binv = cinv(b);
      auto pcr = [&](cx r) {
#pragma unroll
        for (int s = 0; s < 5; s++) {
          const int d = 1 << s;
          cx rm = shfl_up_c(r, d), rp = shfl_dn_c(r, d);
          r = pfma(pa_[s], rm, pfma(pg_[s], rp, r));
        }
        return r * binv;
      };
      const cx GL = pcr(A0), GR = pcr(C30);
      {   // block B: su, (A_off, 0), pa[0..4], pg[0..4];  block C: binv, GL, GR, gl[0..6], gr[0..6]
        cx blk[12];
        blk[0] = su; blk[1] = mk(sAd, 0.0);
#pragma unroll
        for (int s = 0; s < 5; s++) { blk[2 + s] = pa_[s]; blk[7 + s] = pg_[s]; }
        tm_store_cx(tb + TM8_B, blk);
        cx blc[17];
        blc[0] = binv; blc[1] = GL; blc[2] = GR;
#pragma unroll
        for (int k = 0; k < 7; k++) { blc[3 + k] = gl[k]; blc[10 + k] = gr[k]; }
        tm_store_cx(tb + TM8_C, blc);
        tm_wait_st();
      }
      if (lane == 31) { s_l3[wid][0] = l3P; s_l3[wid][1] = l3D; s_l3[wid][2] = l3N;
                        s_l3[wid][7] = mk((wid + 1 < NW) ? sAd : 0.0); s_l3[wid][8] = su; }
      if (lane == 0) { s_l3[wid][3] = GL; s_l3[wid][4] = GR; }
      if (lane == 30) { s_l3[wid][5] = GL; s_l3[wid][6] = GR; }
      __syncthreads();
      if (t < NW) {
        cx cc[NW], dd[NW];
        cx cprev = mk(0.0), dprev = mk(0.0);
#pragma unroll
        for (int w = 0; w < NW; w++) {
          cx Pw = s_l3[w][0], Dw = s_l3[w][1], Nw = s_l3[w][2];
          cx lo = (w > 0) ? -(Pw * s_l3[w][5]) : mk(0.0);
          cx di = nfma(Pw, s_l3[w][6], Dw);
          if (w + 1 < NW) di = nfma(Nw, s_l3[(w + 1) % NW][3], di);
          cx up = (w + 1 < NW) ? -(Nw * s_l3[(w + 1) % NW][4]) : mk(0.0);
          cx rhs = (w == t) ? mk(1.0) : mk(0.0);
          cx deninv = cinv(nfma(lo, cprev, di));
          cc[w] = up * deninv;
          dd[w] = nfma(lo, dprev, rhs) * deninv;
          cprev = cc[w]; dprev = dd[w];
        }
        cx xn = mk(0.0);
#pragma unroll
        for (int w = NW - 1; w >= 0; w--) { xn = nfma(cc[w], xn, dd[w]); s_minv[w][t] = xn; }
        s_c3[t][0] = s_l3[t][0].re; s_c3[t][1] = s_l3[t][0].im; s_c3[t][2] = s_l3[t][8].re; s_c3[t][3] = s_l3[t][8].im;
        s_c3[t][4] = s_l3[t][2].re; s_c3[t][5] = s_l3[t][2].im; s_c3[t][6] = s_l3[t][7].re;
      }
      __syncthreads();
    }
    const int v3 = lane & (NW - 1);
    double q[C], phi[C];
#pragma unroll
    for (int k = 0; k < C; k++) { q[k] = (t * C + k < P.ni) ? 1.0 : 0.0; phi[k] = 0.0; }
    double XL = (t > 0 && t * C - 1 < P.ni) ? 1.0 : 0.0;
    double qn = ((t + 1) * C < P.ni) ? 1.0 : 0.0;
    double *hb = P.hist + (size_t)(P.store_full ? p : blockIdx.x) * P.hist_stride + 2 * t;
    auto store_slice = [&](double *dst) {
#pragma unroll
      for (int k = 0; k < C; k += 2) *reinterpret_cast<double2 *>(dst + k * T) = make_double2(q[k], q[k + 1]);
    };
    store_slice(hb);
    const bool full = P.store_full != 0;
    const unsigned c3_v = smem_u32(&s_c3[v3][0]);
    const unsigned mw_v = smem_u32(&s_minv[wid][v3]), mm_v = smem_u32(&s_minv[(wid + NW - 1) % NW][v3]);
    const bool l0 = (lane == 0), l30 = (lane == 30), l31 = (lane == 31);
    double *hw = hb;
    const double *hr = hb + (size_t)n * SL;

    // block A as three pieces: a0 = al[1..6], ca[0], ca[1]; a1 = ca[2..6], be[0..2]; a2 = be[3..5], sl
    uint32_t a0[32], a1[32], a2[16];
    tm_ld32(tb + TM8_A, a0); tm_ld32(tb + TM8_A + 32, a1); tm_ld16(tb + TM8_A + 64, a2);
    tm_wait_ld();
    tm_pin(a0); tm_pin(a1); tm_pin(a2);

    auto step = [&](auto ph, const int j) {
      constexpr int PH = decltype(ph)::value;
      hw += SL; hr -= SL;
      double2 qo[C / 2];
      if (PH == 2) {
#pragma unroll
        for (int k = 0; k < C; k += 2) qo[k / 2] = *reinterpret_cast<const double2 *>(hr + k * T);
      }
      uint32_t b0[32], b1[16];
      tm_ld32(tb + TM8_B, b0); tm_ld16(tb + TM8_B + 32, b1);
      // ---- chunk solve with zero separators (block A)
      cx z[CI];
      {
        double tk[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) tk[k] = fma(4.0, q[k], ((k == 0) ? XL : q[k - 1]) + q[k + 1]);
        auto cak = [&](int k) { return k < 2 ? tm_getc(a0, 6 + k) : tm_getc(a1, k - 2); };
        auto bek = [&](int k) { return k < 3 ? tm_getc(a1, 5 + k) : tm_getc(a2, k - 3); };
#if IRK4_TWOSIDED
        // eliminations from both chunk ends towards node 3, on u = y / A_off (real right-hand side)
        const cx y1 = nfma(tm_getc(a0, 0), mk(tk[0]), mk(tk[1])), y5 = nfma(tm_getc(a0, 3), mk(tk[6]), mk(tk[5]));
        const cx y2 = nfma(tm_getc(a0, 1), y1, mk(tk[2])), y4 = nfma(tm_getc(a0, 2), y5, mk(tk[4]));
        const cx y3 = nfma(tm_getc(a0, 5), y4, nfma(tm_getc(a0, 4), y2, mk(tk[3])));
        // substitution outwards from node 3
        z[3] = cak(3) * y3;
        z[2] = nfma(bek(2), z[3], cak(2) * y2);
        z[4] = nfma(bek(3), z[3], cak(4) * y4);
        z[1] = nfma(bek(1), z[2], cak(1) * y1);
        z[5] = nfma(bek(4), z[4], cak(5) * y5);
        z[0] = nfma(bek(0), z[1], cak(0) * mk(tk[0]));
        z[6] = nfma(bek(5), z[5], cak(6) * mk(tk[6]));
#else
        z[0] = mk(tk[0]);
#pragma unroll
        for (int k = 1; k < CI; k++) z[k] = nfma(tm_getc(a0, k - 1), z[k - 1], mk(tk[k]));   // al[k]
        z[CI - 1] = cak(CI - 1) * z[CI - 1];
#pragma unroll
        for (int k = CI - 2; k >= 0; k--) z[k] = nfma(bek(k), z[k + 1], cak(k) * z[k]);
#endif
      }
      const cx sl = tm_getc(a2, 3);
      tm_wait_ld();
      tm_pin(b0); tm_pin(b1);
      const cx su = tm_getc(b0, 0);
      const double sAd = tm_get(b0, 2);
      cx r = mk(sAd * fma(4.0, q[C - 1], q[CI - 1]));
      r = nfma(sl, z[CI - 1], r);
      const cx rsep = r;
      {
        r.re = fma(sAd, qn, r.re);
        cx zfn = shfl_dn_c(z[0], 1);
        r = nfma(su, zfn, r);
      }
      if (l31) r = mk(0.0);
      uint32_t c0[32], c1[32], c2[4];
      tm_ld32(tb + TM8_C, c0); tm_ld32(tb + TM8_C + 32, c1); tm_ld4(tb + TM8_C + 64, c2);
      // ---- level 2: pa[s] = b cx 2+s, pg[s] = b cx 7+s  (b0 holds cx 0..7, b1 cx 8..11)
#define IRK8_CR(S, PA, PG)                                                      \
      {                                                                             \
        cx rm = shfl_up_c(r, 1 << (S)), rp = shfl_dn_c(r, 1 << (S));                \
        r = pfma((PA), rm, pfma((PG), rp, r));                                      \
      }
      IRK8_CR(0, tm_getc(b0, 2), tm_getc(b0, 7)) IRK8_CR(1, tm_getc(b0, 3), tm_getc(b1, 0)) IRK8_CR(2, tm_getc(b0, 4), tm_getc(b1, 1))
      IRK8_CR(3, tm_getc(b0, 5), tm_getc(b1, 2)) IRK8_CR(4, tm_getc(b0, 6), tm_getc(b1, 3))
#undef IRK8_CR
      tm_wait_ld();
      tm_pin(c0); tm_pin(c1); tm_pin(c2);
      const cx Z = r * tm_getc(c0, 0);
      double *pb = s_pub[j & 1][wid];
      if (l0) { pb[0] = q[0]; pb[2] = z[0].re; pb[3] = z[0].im; pb[4] = Z.re; pb[5] = Z.im; }
      if (l30) { pb[6] = Z.re; pb[7] = Z.im; }
      if (l31) { pb[8] = rsep.re; pb[9] = rsep.im; }
      __syncthreads();
      tm_ld32(tb + TM8_A, a0); tm_ld32(tb + TM8_A + 32, a1); tm_ld16(tb + TM8_A + 64, a2);   // next step's sweep coefficients
      cx Wm, Ww;
      {
        const double *pv = s_pub[j & 1][v3], *pn = s_pub[j & 1][(v3 + 1) % NW];
        const double2 c01 = lds128(c3_v), c23 = lds128(c3_v + 16), c45 = lds128(c3_v + 32);
        const double c3Au = lds64(c3_v + 48);
        const double2 mw = lds128(mw_v), mm = lds128(mm_v);
        cx R = nfma(mk(c01.x, c01.y), mk(pv[6], pv[7]), mk(pv[8], pv[9]));
        cx R2 = nfma(mk(c23.x, c23.y), mk(pn[2], pn[3]), mk(c3Au * pn[0]));
        R2 = nfma(mk(c45.x, c45.y), mk(pn[4], pn[5]), R2);
        R = R + R2;
        Ww = mk(mw.x, mw.y) * R;
        Wm = (wid > 0) ? mk(mm.x, mm.y) * R : mk(0.0);
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
          Ww.re += __shfl_xor_sync(0xffffffffu, Ww.re, d); Ww.im += __shfl_xor_sync(0xffffffffu, Ww.im, d);
          Wm.re += __shfl_xor_sync(0xffffffffu, Wm.re, d); Wm.im += __shfl_xor_sync(0xffffffffu, Wm.im, d);
        }
      }
      const cx Y = l31 ? Ww : nfma(tm_getc(c0, 1), Wm, nfma(tm_getc(c0, 2), Ww, Z));
      cx YL = shfl_up_c(Y, 1);
      if (l0) YL = Wm;
      // q+ = q - 2 Re[alpha y]: gl[k] = c cx 3+k, gr[k] = c cx 10+k  (c0 holds cx 0..7, c1 cx 8..15, c2 cx 16)
      {
        auto glk = [&](int k) { return k < 5 ? tm_getc(c0, 3 + k) : tm_getc(c1, k - 5); };
        auto grk = [&](int k) { return k < 6 ? tm_getc(c1, 2 + k) : mk(tm_get(c2, 0), tm_get(c2, 1)); };
#pragma unroll
        for (int k = 0; k < CI; k++) {
          const cx yk = nfma(glk(k), YL, nfma(grk(k), Y, z[k]));
          q[k] = fma(-a_re, yk.re, fma(-a_im, yk.im, q[k]));
        }
      }
      q[C - 1] = fma(-a_re, Y.re, fma(-a_im, Y.im, q[C - 1]));
      XL = fma(-a_re, YL.re, fma(-a_im, YL.im, XL));
      qn = shfl_dn_d(q[0], 1);
      if (l31) qn = 0.0;
      if (PH == 0 || full) store_slice(hw);
      if (PH >= 1) {
        const double wj = __ldg(P.w + j);
#pragma unroll
        for (int k = 0; k < C; k += 2) {
          phi[k] = fma(wj * q[k], PH == 2 ? qo[k / 2].x : q[k], phi[k]);
          phi[k + 1] = fma(wj * q[k + 1], PH == 2 ? qo[k / 2].y : q[k + 1], phi[k + 1]);
        }
      }
      tm_wait_ld();
      tm_pin(a0); tm_pin(a1); tm_pin(a2);
    };
    int j = 1;
    for (; 2 * j < n; j++) step(std::integral_constant<int, 0>{}, j);
    if (2 * j == n) { step(std::integral_constant<int, 1>{}, j); j++; }
    for (; j <= n; j++) step(std::integral_constant<int, 2>{}, j);

    double qsum = 0.0;
    const double hcell = L / (P.N - 1);
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int g = t * C + k;
      if (g < P.ni) {
        const int i = g + 1;
        const double f0 = P.f0[(size_t)pp * P.N + i];
        P.out[(size_t)p * P.out_stride + g] = P.sign * (f0 - phi[k]);
        if (!P.pshare) P.phi[(size_t)p * P.N + i] = phi[k];
        qsum += (0.5 * (hcell + hcell)) * q[k];
        if (P.eta_full && !P.pshare) P.eta_full[(size_t)p * P.N + i] = P.eta_mid[(size_t)p * P.eta_stride + g];
      }
    }
    if (t == 0 && !P.pshare) {
      P.phi[(size_t)p * P.N] = 0.0; P.phi[(size_t)p * P.N + P.N - 1] = 0.0;
      if (P.eta_full) {
        P.eta_full[(size_t)p * P.N] = eta_node(P, p, 0, L);
        P.eta_full[(size_t)p * P.N + P.N - 1] = eta_node(P, p, P.N - 1, L);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, d);
    if (lane == 0) s_red[wid] = qsum;
    __syncthreads();
    if (t == 0 && !P.pshare) {
      double s = 0.0;
      for (int w = 0; w < NW; w++) s += s_red[w];
      P.Q[p] = s / L;
    }
  }
  __syncthreads();
  if (wid == 0) tm4_dealloc(s_tm);
}

}  // namespace scftb
