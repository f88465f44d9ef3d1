// march_irk4_tm.cuh — the IRK4 contour march (march_irk4.cuh: the reference's 2-stage Gauss-Legendre stepper as ONE complex
// symmetric tridiagonal solve per step; scft.cc:671-693, drivescft.cc:130-146) for the benchmarked shape — uniform mesh,
// 513..1024 unknowns, C = 4 nodes per thread, T = 256 threads — with its loop-invariant complex coefficients in TENSOR MEMORY.
//
// Same algorithm and arithmetic as march_irk4_kernel<4,256,true>.  That kernel needs 218 registers (one CTA of 8 warps per
// SM) and runs at 34 % of the fp64 pipe, latency-bound.  Here each thread keeps 56 doubles in its TMEM lane (warps 0-3 in
// columns [0,112), warps 4-7 — which map onto the same 128 lanes — in [112,224)) and streams them back with tcgen05.ld one
// phase ahead of use; the eight per-warp level-3 constants come from shared memory.  <= 128 registers, 2 CTAs = 16 warps per
// SM (2 x 256 columns = the whole tensor memory).  Measured: 1.52e11 -> 1.89e11 DOF-steps/s on the 4096-problem sweep; with 16
// warps the kernel is ISSUE-bound (407 SASS instructions per warp-step for 128 nodes, 74 of them 32-bit shuffles of the
// complex cyclic reduction; ncu: issue slots 48 %, fp64 pipe 41 %), so the next lever is C = 8 nodes per thread, not
// occupancy.  A variant that keeps the chunk-sweep coefficients in registers and loads everything else at the top of the
// step spills more and is 3 % slower.
//
// TMEM column map of a thread (a double is two 32-bit columns):
//   block A  [  0, 32)  al1, al2, ca0, ca1, ca2, be0, be1, sl          chunk sweeps, separator row       (complex)
//   block B1 [ 32, 40)  su (complex), A_off, -                         separator row
//   block B2 [ 40, 80)  pa[0..4], pg[0..4]                             cyclic reduction                  (complex)
//   block C  [ 80,112)  binv, GL, GR, gl0, gl1, gl2, gr0, gr1          back substitution                 (complex; gr2 stays in registers)
#pragma once
#include "march1d_tmem.cuh"
#include "march_irk4.cuh"

namespace scftb {

constexpr int TM4_COLS = 256, TM4_PER = 112;
constexpr int TM4_A = 0, TM4_B1 = 32, TM4_B2 = 40, TM4_C = 80;

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

__global__ void __launch_bounds__(256, 2) march_irk4_tm_kernel(MarchParams P) {
  constexpr int C = 4, T = 256, CI = 3, NW = 8, SL = T * C;
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
  const double a_re = 12.0, a_im = 12.0 * 1.7320508075688772;   // 2*Re[alpha y] = 12 y_re + 12 sqrt3 y_im

  if (wid == 0) tm4_alloc(&s_tm);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  // this thread's lane (warp w reaches lanes 32 (w % 4) ...) and its first column
  const uint32_t tb = s_tm + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)((wid >> 2) * TM4_PER);

  for (int p = blockIdx.x; p < P.nprob; p += gridDim.x) {
    if (P.skip && P.skip[p]) continue;
    const int pp = P.pshare ? 0 : p;
    const double L = P.L[pp];
    cx gr2;
    {
      auto wrow = [&](const Row &r, cx &wl, cx &wd, cx &wu) {   // W = z1 A + dt D
        wl = mk(fma(dt, r.Dl, z1.re * r.Al), z1.im * r.Al);
        wd = mk(fma(dt, r.Dd, z1.re * r.Ad), z1.im * r.Ad);
        wu = mk(fma(dt, r.Du, z1.re * r.Au), z1.im * r.Au);
      };
      // ---------------------------------------------------------------- assembly + level 1 (march_irk4.cuh, UNI)
      cx ca[CI], al[CI], be[CI], gl[CI], gr[CI];
      double sAd;
      cx sl, sd, su;
      {
        Row rs = assemble_row(P, p, t * C + CI, L, dt);
        sAd = (rs.Al != 0.0) ? rs.Al : rs.Au;   // A's row is A_off (1,4,1)
        if (t * C + CI >= P.ni) { sl = mk(0.0); sd = mk(1.0); su = mk(0.0); }   // padding: identity row
        else wrow(rs, sl, sd, su);
      }
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
      gr2 = gr[2];
      {   // blocks A and B1 are final
        uint32_t w32[32];
        tm_putc(w32, 0, al[1]); tm_putc(w32, 1, al[2]); tm_putc(w32, 2, ca[0]); tm_putc(w32, 3, ca[1]); tm_putc(w32, 4, ca[2]);
        tm_putc(w32, 5, be[0]); tm_putc(w32, 6, be[1]); tm_putc(w32, 7, sl);
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int i = 0; i < 16; i++) { lo[i] = w32[i]; hi[i] = w32[16 + i]; }
          tm_st16(tb + TM4_A, lo); tm_st16(tb + TM4_A + 16, hi);
        }
        uint32_t w8[8];
        tm_put(w8, 0, su.re); tm_put(w8, 1, su.im); tm_put(w8, 2, sAd); tm_put(w8, 3, 0.0);
        tm_st8(tb + TM4_B1, w8);
      }
      // ---------------------------------------------------------------- Schur rows on the separators
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
      // ---------------------------------------------------------------- level 2: cyclic reduction
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
      {   // blocks B2 and C
        uint32_t w32[32], w8[8];
#pragma unroll
        for (int s = 0; s < 5; s++) tm_putc(w32, s, pa_[s]);
#pragma unroll
        for (int s = 0; s < 3; s++) tm_putc(w32, 5 + s, pg_[s]);
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int i = 0; i < 16; i++) { lo[i] = w32[i]; hi[i] = w32[16 + i]; }
          tm_st16(tb + TM4_B2, lo); tm_st16(tb + TM4_B2 + 16, hi);
        }
        tm_putc(w8, 0, pg_[3]); tm_putc(w8, 1, pg_[4]);
        tm_st8(tb + TM4_B2 + 32, w8);
        tm_putc(w32, 0, binv); tm_putc(w32, 1, GL); tm_putc(w32, 2, GR); tm_putc(w32, 3, gl[0]); tm_putc(w32, 4, gl[1]); tm_putc(w32, 5, gl[2]);
        tm_putc(w32, 6, gr[0]); tm_putc(w32, 7, gr[1]);
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int i = 0; i < 16; i++) { lo[i] = w32[i]; hi[i] = w32[16 + i]; }
          tm_st16(tb + TM4_C, lo); tm_st16(tb + TM4_C + 16, hi);
        }
        tm_wait_st();
      }
      // ---------------------------------------------------------------- level 3 setup
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
        // the constants lane v3 needs to form R_v3, in a conflict-free table
        s_c3[t][0] = s_l3[t][0].re; s_c3[t][1] = s_l3[t][0].im; s_c3[t][2] = s_l3[t][8].re; s_c3[t][3] = s_l3[t][8].im;
        s_c3[t][4] = s_l3[t][2].re; s_c3[t][5] = s_l3[t][2].im; s_c3[t][6] = s_l3[t][7].re;
      }
      __syncthreads();
    }
    const int v3 = lane & (NW - 1);
    // ---------------------------------------------------------------- initial condition
    double q[C], phi[C];
#pragma unroll
    for (int k = 0; k < C; k++) { q[k] = (t * C + k < P.ni) ? 1.0 : 0.0; phi[k] = 0.0; }
    double XL = (t > 0 && t * C - 1 < P.ni) ? 1.0 : 0.0;
    double qn = ((t + 1) * C < P.ni) ? 1.0 : 0.0;
    double *hb = P.hist + (size_t)(P.store_full ? p : blockIdx.x) * P.hist_stride + 2 * t;   // this thread's pair (k, k+1) of a slice
    auto store_slice = [&](double *dst) {
      *reinterpret_cast<double2 *>(dst) = make_double2(q[0], q[1]);
      *reinterpret_cast<double2 *>(dst + 2 * T) = make_double2(q[2], q[3]);
    };
    store_slice(hb);
    const bool full = P.store_full != 0;
    const unsigned c3_v = smem_u32(&s_c3[v3][0]);
    const unsigned mw_v = smem_u32(&s_minv[wid][v3]), mm_v = smem_u32(&s_minv[(wid + NW - 1) % NW][v3]);


    uint32_t rA[32];
    tm_ld32(tb + TM4_A, rA);
    tm_wait_ld();
    tm_pin(rA);

    // ---------------------------------------------------------------- the contour march
    // PH = 0: 2j < n (store the slice), 1: 2j = n (pairs with itself), 2: 2j > n (pairs with slice n-j); one instantiation
    // per phase keeps the selects and branches of the quadrature out of the step
    const bool l0 = (lane == 0), l30 = (lane == 30), l31 = (lane == 31);
    double *hw = hb;                              // write cursor (slice j)
    const double *hr = hb + (size_t)n * SL;       // read cursor (slice n-j)
    auto step = [&](auto ph, const int j) {
      constexpr int PH = decltype(ph)::value;
      hw += SL; hr -= SL;
      double2 qo01 = make_double2(0.0, 0.0), qo23 = make_double2(0.0, 0.0);
      if (PH == 2) {
        qo01 = *reinterpret_cast<const double2 *>(hr);
        qo23 = *reinterpret_cast<const double2 *>(hr + 2 * T);
      }
      uint32_t rB1[8];
      tm_ld8(tb + TM4_B1, rB1);
      // ---- chunk solve with zero separators (block A): forward on u = y / A_off (real right-hand side), backward
      cx z[CI];
      {
        const double t0 = fma(4.0, q[0], XL + q[1]), t1 = fma(4.0, q[1], q[0] + q[2]), t2 = fma(4.0, q[2], q[1] + q[3]);
        z[0] = mk(t0);
        z[1] = nfma(tm_getc(rA, 0), z[0], mk(t1));
        z[2] = nfma(tm_getc(rA, 1), z[1], mk(t2));
        z[2] = tm_getc(rA, 4) * z[2];
        z[1] = nfma(tm_getc(rA, 6), z[2], tm_getc(rA, 3) * z[1]);
        z[0] = nfma(tm_getc(rA, 5), z[1], tm_getc(rA, 2) * z[0]);
      }
      const cx sl = tm_getc(rA, 7);
      uint32_t rB2[32], rB3[8];
      tm_ld32(tb + TM4_B2, rB2);
      tm_ld8(tb + TM4_B2 + 32, rB3);
      tm_wait_ld();
      tm_pin(rB1); tm_pin(rB2); tm_pin(rB3);
      const cx su = mk(tm_get(rB1, 0), tm_get(rB1, 1));
      const double sAd = tm_get(rB1, 2);
      cx r = mk(sAd * fma(4.0, q[C - 1], q[CI - 1]));
      r = nfma(sl, z[CI - 1], r);
      const cx rsep = r;
      {
        r.re = fma(sAd, qn, r.re);
        cx zfn = shfl_dn_c(z[0], 1);
        r = nfma(su, zfn, r);
      }
      if (l31) r = mk(0.0);
      // ---- level 2 (block B2): five cyclic-reduction stages
#define IRK4_CR_STAGE(S, PG)                                                    \
      {                                                                             \
        cx rm = shfl_up_c(r, 1 << (S)), rp = shfl_dn_c(r, 1 << (S));                \
        r = pfma(tm_getc(rB2, (S)), rm, pfma((PG), rp, r));                         \
      }
      IRK4_CR_STAGE(0, tm_getc(rB2, 5)) IRK4_CR_STAGE(1, tm_getc(rB2, 6)) IRK4_CR_STAGE(2, tm_getc(rB2, 7))
      IRK4_CR_STAGE(3, tm_getc(rB3, 0)) IRK4_CR_STAGE(4, tm_getc(rB3, 1))
#undef IRK4_CR_STAGE
      uint32_t rC[32];
      tm_ld32(tb + TM4_C, rC);
      tm_ld32(tb + TM4_A, rA);      // next step's sweep coefficients
      tm_wait_ld();
      tm_pin(rC); tm_pin(rA);
      const cx Z = r * tm_getc(rC, 0);
      double *pb = s_pub[j & 1][wid];
      if (l0) { pb[0] = q[0]; pb[2] = z[0].re; pb[3] = z[0].im; pb[4] = Z.re; pb[5] = Z.im; }
      if (l30) { pb[6] = Z.re; pb[7] = Z.im; }
      if (l31) { pb[8] = rsep.re; pb[9] = rsep.im; }
      __syncthreads();
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
      const cx Y = l31 ? Ww : nfma(tm_getc(rC, 1), Wm, nfma(tm_getc(rC, 2), Ww, Z));   // solution at the own separator
      cx YL = shfl_up_c(Y, 1);
      if (l0) YL = Wm;                                                                 // previous warp's separator
      // q+ = q - 2 Re[alpha y]
      {
        cx y0 = nfma(tm_getc(rC, 3), YL, nfma(tm_getc(rC, 6), Y, z[0]));
        cx y1 = nfma(tm_getc(rC, 4), YL, nfma(tm_getc(rC, 7), Y, z[1]));
        cx y2 = nfma(tm_getc(rC, 5), YL, nfma(gr2, Y, z[2]));
        q[0] = fma(-a_re, y0.re, fma(-a_im, y0.im, q[0]));
        q[1] = fma(-a_re, y1.re, fma(-a_im, y1.im, q[1]));
        q[2] = fma(-a_re, y2.re, fma(-a_im, y2.im, q[2]));
      }
      q[C - 1] = fma(-a_re, Y.re, fma(-a_im, Y.im, q[C - 1]));
      XL = fma(-a_re, YL.re, fma(-a_im, YL.im, XL));
      qn = shfl_dn_d(q[0], 1);
      if (l31) qn = 0.0;
      if (PH == 0 || full) store_slice(hw);
      if (PH >= 1) {
        const double wj = __ldg(P.w + j);
        phi[0] = fma(wj * q[0], PH == 2 ? qo01.x : q[0], phi[0]);
        phi[1] = fma(wj * q[1], PH == 2 ? qo01.y : q[1], phi[1]);
        phi[2] = fma(wj * q[2], PH == 2 ? qo23.x : q[2], phi[2]);
        phi[3] = fma(wj * q[3], PH == 2 ? qo23.y : q[3], phi[3]);
      }
    };
    int j = 1;
    // (two steps per loop trip, as in march_tm_kernel, spill 780 B here and are 1.5 % slower)
    for (; 2 * j < n; j++) step(std::integral_constant<int, 0>{}, j);
    if (2 * j == n) { step(std::integral_constant<int, 1>{}, j); j++; }
    for (; j <= n; j++) step(std::integral_constant<int, 2>{}, j);

    // ---------------------------------------------------------------- residual, phi, Q
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
