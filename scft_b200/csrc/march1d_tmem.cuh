// march1d_tmem.cuh — the contour-march kernel for the benchmarked shape (uniform mesh, 513 < unknowns <= 1024, even
// number of steps, one species) with its per-thread solver coefficients held in TENSOR MEMORY.
//
// Same algorithm and arithmetic as march_ie_kernel<8,128,true,...> (march1d.cuh: three-level substructured solve of the
// constant tridiagonal system, half history, fused quadrature; reference path 1D_FEM.c:95-228, drivescft.cc:130-213).
// What changes is where the loop-invariant coefficients live.  The register-resident kernel needs 168 registers per
// thread (35 chunk coefficients + 23 separator / warp / CTA-level constants + q, phi), which caps an SM at 3 CTAs = 12
// warps, and the kernel is latency-bound there (ncu: 18.7 % warps active, 35 % stall_wait, 33 % short scoreboard).
// Blackwell's tensor memory (256 KB per SM, 512 columns x 128 lanes x 32 bit) is idle in an fp64 kernel and can be read
// with tcgen05.ld at > 600 B/clk/SM (tools/tmem_bw.cu: 5x shared memory, ~43 cycles round trip), so it serves as a
// second register file: each thread owns one TMEM lane and keeps 56 doubles of coefficients in 112 columns, streamed into
// registers one phase ahead of their use.  The thread keeps q, phi and the block in flight: <= 128 registers, 4 CTAs per
// SM (the whole TMEM: 4 x 128 columns), 16 warps.
//
// TMEM column map of a thread (a double is two 32-bit columns):
//   block A  [  0, 16)  al[0..5], ca[0], sl                                        first sweep
//   block B  [ 16, 48)  ca[1..6], be[1..6], A_off, su                              second sweep, separator row
//   block C  [ 48, 80)  pa[0..2], pg[0..2], inv4[0..3], kP, kAu, ksu, kNx, mW, mM  warp and CTA level
//   block D  [ 80,112)  gl[0..6], gr[0..6], GL, GR                                 spikes
#pragma once
#include <type_traits>

#include "march1d.cuh"

namespace scftb {

#ifndef TM_TWOSIDED
// 1: two-sided chunk elimination (dependent chain per step 8 instead of 13 operations; single-CTA latency 989 -> 848
// cycles per step) — measured NOT to raise the throughput at four CTAs per SM (1740 vs 1673 cycles per four steps: its
// extra temporaries spill), so the one-sided UL sweeps stay the default
#define TM_TWOSIDED 0
#endif
#ifndef TM_PREFETCH
#define TM_PREFETCH 2   // contour steps the paired history slice is fetched ahead (ring of 4 staging buffers)
#endif
constexpr int TM_COLS = 128;
constexpr int TM_A = 0, TM_B = 16, TM_C = 48, TM_D = 80;

__device__ __forceinline__ void tm_alloc(uint32_t *smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(TM_COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TM_COLS) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// the loaded registers are defined only after tcgen05.wait::ld: make every later use depend on an (empty) asm statement
// that follows the wait in program order, so the compiler cannot hoist arithmetic on them above it
template <int NR>
__device__ __forceinline__ void tm_pin(uint32_t (&r)[NR]) {
#pragma unroll
  for (int i = 0; i < NR; i++) asm volatile("" : "+r"(r[i]));
}
template <int NR>
__device__ __forceinline__ double tm_get(const uint32_t (&r)[NR], int i) { return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]); }
template <int NR>
__device__ __forceinline__ void tm_put(uint32_t (&r)[NR], int i, double v) {
  r[2 * i] = (uint32_t)__double2loint(v);
  r[2 * i + 1] = (uint32_t)__double2hiint(v);
}

// C = 8 nodes per thread, T = 128 threads per problem, uniform mesh, even nsteps, one species.
__global__ void __launch_bounds__(128, 4) march_tm_kernel(MarchParams P) {
  constexpr int C = 8, T = 128, CI = 7, NW = 4, SL = T * C, NST = 3;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  __shared__ __align__(16) double s_ex[2][T];        // setup exchange gl0 / gr0
  __shared__ __align__(16) double s_l3[NW][4];       // per warp separator v: P, A_off, su, Nx
  __shared__ __align__(16) double s_l3s[NW][6];      // setup only: D, GL0, GR0, GL30, GR30
  __shared__ __align__(16) double s_minv[NW][NW];    // inverse of the level-3 matrix
  __shared__ __align__(16) double s_pub[2][NW][2];   // per step: {A_v, B_{v+1}} -> R_v = A_v + B_{v+1} (double-buffered)
  __shared__ double s_red[NW];
  // staging ring for the paired history slices q(., n-j): fetched PD steps ahead, so that PD slices per CTA (x 4 CTAs per
  // SM) are in flight — one slice per CTA is not enough memory-level parallelism to cover the loaded HBM latency
  constexpr int RING = 4, PD = TM_PREFETCH;
  static_assert(PD >= 1 && PD < RING, "prefetch distance");
  __shared__ __align__(16) double s_qo[RING][C / 2][T][2];
  __shared__ uint32_t s_tm;
  // keep at most four CTAs on an SM whatever the register count turns out to be: the fifth could not allocate its
  // tensor-memory columns (4 x 128 = all 512) and would spin in tcgen05.alloc
  extern __shared__ double s_pad[];
  const int n = P.nsteps;
  const double dt = 1.0 / n;

  if (wid == 0) tm_alloc(&s_tm);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t tb = s_tm + ((uint32_t)(wid * 32) << 16);   // this thread's lane, column 0

  for (int p = blockIdx.x; p < P.nprob; p += gridDim.x) {
    if (P.skip && P.skip[p]) continue;
    const int pp = P.pshare ? 0 : p;
    const double L = P.L[pp];
    double q[C], phi[C], XL, qn, GLm, GRm;
    {
      // ---------------------------------------------------------------- assembly + level 1 (see march1d.cuh)
      double ca[CI], al[CI], be[CI], gl[CI], gr[CI];
      double sAd, sl, sd, su;
      {
        Row rs = assemble_row(P, p, t * C + CI, L, dt);
        sl = rs.Tl; sd = rs.Td; su = rs.Tu;
        sAd = (rs.Al != 0.0) ? rs.Al : rs.Au;    // A_off; Dirichlet zeroing is carried by the neighbour values
      }
#if TM_TWOSIDED
      // Two-sided ("burn at both ends") elimination of the chunk: nodes 0..2 are eliminated downwards, nodes 6..4 upwards,
      // both meet at node 3, and the substitution runs outwards from there.  Same operation count as the one-sided sweeps,
      // but the dependent chain per contour step is 2 + 2 + 1 + 3 operations instead of 6 + 1 + 6.
      //   al[] <- the six elimination multipliers  m1, m2, m4, m5, mL3, mR3        (block A)
      //   ca[] <- A_off / pivot_k  (k = 0..6; node 3 goes to block A)               (block B)
      //   be[] <- substitution couplings d0, d1, d2 (to node k+1), d4, d5, d6 (to node k-1), be[0] unused
      {
        Row rw[CI];
        double piv[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) rw[k] = assemble_row(P, p, t * C + k, L, dt);
        double mL[CI], mR[CI];   // multipliers applied to the upper / lower neighbour's row
        piv[0] = rw[0].Td;
#pragma unroll
        for (int k = 1; k <= 2; k++) { mL[k] = rw[k].Tl / piv[k - 1]; piv[k] = rw[k].Td - mL[k] * rw[k - 1].Tu; }
        piv[6] = rw[6].Td;
#pragma unroll
        for (int k = 5; k >= 4; k--) { mR[k] = rw[k].Tu / piv[k + 1]; piv[k] = rw[k].Td - mR[k] * rw[k + 1].Tl; }
        mL[3] = rw[3].Tl / piv[2]; mR[3] = rw[3].Tu / piv[4];
        piv[3] = rw[3].Td - mL[3] * rw[2].Tu - mR[3] * rw[4].Tl;
        auto solve = [&](const double (&rhs)[CI], double (&x)[CI]) {   // exact chunk solve with these factors
          double y[CI];
          y[0] = rhs[0]; y[1] = rhs[1] - mL[1] * y[0]; y[2] = rhs[2] - mL[2] * y[1];
          y[6] = rhs[6]; y[5] = rhs[5] - mR[5] * y[6]; y[4] = rhs[4] - mR[4] * y[5];
          y[3] = rhs[3] - mL[3] * y[2] - mR[3] * y[4];
          x[3] = y[3] / piv[3];
          x[2] = (y[2] - rw[2].Tu * x[3]) / piv[2]; x[1] = (y[1] - rw[1].Tu * x[2]) / piv[1]; x[0] = (y[0] - rw[0].Tu * x[1]) / piv[0];
          x[4] = (y[4] - rw[4].Tl * x[3]) / piv[4]; x[5] = (y[5] - rw[5].Tl * x[4]) / piv[5]; x[6] = (y[6] - rw[6].Tl * x[5]) / piv[6];
        };
        {   // spikes: T_loc gl = Tl_first e_first, T_loc gr = Tu_last e_last
          double e[CI] = {rw[0].Tl, 0, 0, 0, 0, 0, 0};
          solve(e, gl);
          double f[CI] = {0, 0, 0, 0, 0, 0, rw[6].Tu};
          solve(f, gr);
        }
        al[0] = mL[1]; al[1] = mL[2]; al[2] = mR[4]; al[3] = mR[5]; al[4] = mL[3]; al[5] = mR[3]; al[6] = 0.0;
#pragma unroll
        for (int k = 0; k < CI; k++) {
          const double aoff = (rw[k].Al == 0.0) ? rw[k].Au : rw[k].Al;
          ca[k] = aoff / piv[k];
          be[k] = (k < 3) ? rw[k].Tu / piv[k] : ((k > 3) ? rw[k].Tl / piv[k] : 0.0);
        }
      }
#else
      {
        double Tl0 = 0, TuL = 0, pinv_next = 0, Tl_next = 0;
        double ulo[CI];
#pragma unroll
        for (int k = CI - 1; k >= 0; k--) {
          Row r = assemble_row(P, p, t * C + k, L, dt);
          double piv = (k == CI - 1) ? r.Td : r.Td - (r.Tu * pinv_next) * Tl_next;
          double pinv = 1.0 / piv;
          ca[k] = pinv * ((r.Al == 0.0) ? r.Au : r.Al);
          ulo[k] = (k == CI - 1) ? 0.0 : pinv * r.Tu;
          al[k] = (k == CI - 1) ? 0.0 : r.Tu * pinv_next;
          be[k] = (k == 0) ? 0.0 : pinv * r.Tl;
          if (k == 0) Tl0 = pinv * r.Tl;
          if (k == CI - 1) TuL = pinv * r.Tu;
          pinv_next = pinv; Tl_next = r.Tl;
        }
        gl[0] = Tl0;
#pragma unroll
        for (int k = 1; k < CI; k++) gl[k] = -be[k] * gl[k - 1];
        double y[CI];
        y[CI - 1] = TuL;
#pragma unroll
        for (int k = CI - 2; k >= 0; k--) y[k] = -ulo[k] * y[k + 1];
        gr[0] = y[0];
#pragma unroll
        for (int k = 1; k < CI; k++) gr[k] = y[k] - be[k] * gr[k - 1];
      }
#endif
      {   // blocks A, B and the spikes of block D are final: park them in tensor memory
        uint32_t w16[16];
#if TM_TWOSIDED
        // block A: m1, m2, m4, m5, mL3, mR3, c3, sl     block B: c0,c1,c2,c4,c5,c6, d0,d1,d2,d4,d5,d6, A_off, su
#pragma unroll
        for (int k = 0; k < 6; k++) tm_put(w16, k, al[k]);
        tm_put(w16, 6, ca[3]); tm_put(w16, 7, sl);
        tm_st16(tb + TM_A, w16);
        tm_put(w16, 0, ca[0]); tm_put(w16, 1, ca[1]); tm_put(w16, 2, ca[2]); tm_put(w16, 3, ca[4]); tm_put(w16, 4, ca[5]); tm_put(w16, 5, ca[6]);
        tm_put(w16, 6, be[0]); tm_put(w16, 7, be[1]);
        tm_st16(tb + TM_B, w16);
        tm_put(w16, 0, be[2]); tm_put(w16, 1, be[4]); tm_put(w16, 2, be[5]); tm_put(w16, 3, be[6]);
        tm_put(w16, 4, sAd); tm_put(w16, 5, su); tm_put(w16, 6, 0.0); tm_put(w16, 7, 0.0);
        tm_st16(tb + TM_B + 16, w16);
#else
#pragma unroll
        for (int k = 0; k < 6; k++) tm_put(w16, k, al[k]);
        tm_put(w16, 6, ca[0]); tm_put(w16, 7, sl);
        tm_st16(tb + TM_A, w16);
#pragma unroll
        for (int k = 0; k < 6; k++) tm_put(w16, k, ca[k + 1]);
        tm_put(w16, 6, be[1]); tm_put(w16, 7, be[2]);
        tm_st16(tb + TM_B, w16);
#pragma unroll
        for (int k = 0; k < 4; k++) tm_put(w16, k, be[k + 3]);
        tm_put(w16, 4, sAd); tm_put(w16, 5, su); tm_put(w16, 6, 0.0); tm_put(w16, 7, 0.0);
        tm_st16(tb + TM_B + 16, w16);
#endif
#pragma unroll
        for (int k = 0; k < 7; k++) tm_put(w16, k, gl[k]);
        tm_put(w16, 7, gr[0]);
        tm_st16(tb + TM_D, w16);
      }
      // ---------------------------------------------------------------- Schur rows on separators
      double a, b, c;
      __syncthreads();  // previous problem's readers of shared memory are done
      s_ex[0][t] = gl[0]; s_ex[1][t] = gr[0];
      __syncthreads();
      {
        double gl0n = (t + 1 < T) ? s_ex[0][t + 1] : 0.0, gr0n = (t + 1 < T) ? s_ex[1][t + 1] : 0.0;
        a = -sl * gl[CI - 1];
        b = sd - sl * gr[CI - 1] - su * gl0n;
        c = -su * gr0n;
      }
      // ---------------------------------------------------------------- level 2: warp cyclic reduction setup
      const double l3P = a, l3D = b, l3N = c;
      const double A0 = (lane == 0) ? a : 0.0, C30 = (lane == 30) ? c : 0.0;
      if (lane == 31) { a = 0.0; b = 1.0; c = 0.0; }
      if (lane == 0) a = 0.0;
      if (lane == 30) c = 0.0;
      double pa_[NST], pg_[NST];
#pragma unroll
      for (int s = 0; s < NST; s++) {
        const int d = 1 << s;
        double am = shfl_up_d(a, d), bm = shfl_up_d(b, d), cm = shfl_up_d(c, d);
        double ap = shfl_dn_d(a, d), bp = shfl_dn_d(b, d), cp = shfl_dn_d(c, d);
        double alpha = (lane >= d) ? -a / bm : 0.0;
        double gamma = (lane + d <= 31) ? -c / bp : 0.0;
        if (lane < d) { am = 0.0; cm = 0.0; }
        if (lane + d > 31) { ap = 0.0; cp = 0.0; }
        b = b + alpha * cm + gamma * ap;
        a = alpha * am;
        c = gamma * cp;
        pa_[s] = alpha; pg_[s] = gamma;
      }
      double inv4[4];
      {
        const int g8 = lane & 7, m8 = lane >> 3;
        double ra[4], rb[4], rc[4];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) { ra[mm] = shfl_d(a, g8 + 8 * mm); rb[mm] = shfl_d(b, g8 + 8 * mm); rc[mm] = shfl_d(c, g8 + 8 * mm); }
        double cpv[4], dpv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double lo = (i > 0) ? rc[i - 1] : 0.0, up = (i < 3) ? ra[i + 1] : 0.0, rhs = (i == m8) ? 1.0 : 0.0;
          double den = rb[i] - lo * ((i > 0) ? cpv[i > 0 ? i - 1 : 0] : 0.0);
          cpv[i] = up / den;
          dpv[i] = (rhs - lo * ((i > 0) ? dpv[i > 0 ? i - 1 : 0] : 0.0)) / den;
        }
        inv4[3] = dpv[3];
#pragma unroll
        for (int i = 2; i >= 0; i--) inv4[i] = dpv[i] - cpv[i] * inv4[i + 1];
      }
      auto pcr = [&](double r) {
#pragma unroll
        for (int s = 0; s < NST; s++) {
          const int d = 1 << s;
          double rm = shfl_up_d(r, d), rp = shfl_dn_d(r, d);
          r = fma(pa_[s], rm, fma(pg_[s], rp, r));
        }
        const int g8 = lane & 7;
        double r0 = shfl_d(r, g8), r1 = shfl_d(r, g8 + 8), r2 = shfl_d(r, g8 + 16), r3 = shfl_d(r, g8 + 24);
        return fma(inv4[0], r0, inv4[1] * r1) + fma(inv4[2], r2, inv4[3] * r3);
      };
      const double GL = pcr(A0), GR = pcr(C30);
      GLm = shfl_up_d(GL, 1); GRm = shfl_up_d(GR, 1);
      if (lane == 0) { GLm = -1.0; GRm = 0.0; }
      // ---------------------------------------------------------------- level 3 setup
      if (lane == 31) {
        s_l3[wid][0] = l3P; s_l3[wid][1] = (wid + 1 < NW) ? sAd : 0.0;
        s_l3[wid][2] = su; s_l3[wid][3] = l3N;
        s_l3s[wid][0] = l3D;
      }
      if (lane == 0) { s_l3s[wid][1] = GL; s_l3s[wid][2] = GR; }
      if (lane == 30) { s_l3s[wid][3] = GL; s_l3s[wid][4] = GR; }
      if (t < 2 * NW) { s_pub[0][t >> 1][t & 1] = 0.0; s_pub[1][t >> 1][t & 1] = 0.0; }   // B_{NW} = 0 is never written again
      __syncthreads();
      if (t < NW) {  // thread v: column v of M^-1 by Thomas
        double cc[NW], dd[NW];
        double cprev = 0.0, dprev = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          double Pw = s_l3[w][0], Dw = s_l3s[w][0], Nw = s_l3[w][3];
          double lo = (w > 0) ? -Pw * s_l3s[w][3] : 0.0;
          double di = Dw - Pw * s_l3s[w][4] - ((w + 1 < NW) ? Nw * s_l3s[(w + 1) % NW][1] : 0.0);
          double up = (w + 1 < NW) ? -Nw * s_l3s[(w + 1) % NW][2] : 0.0;
          double rhs = (w == t) ? 1.0 : 0.0;
          double den = di - lo * cprev;
          cc[w] = up / den;
          dd[w] = (rhs - lo * dprev) / den;
          cprev = cc[w]; dprev = dd[w];
        }
        double xn = 0.0;
#pragma unroll
        for (int w = NW - 1; w >= 0; w--) {
          xn = dd[w] - cc[w] * xn;
          s_minv[w][t] = xn;
        }
      }
      __syncthreads();
      {   // block C and the rest of block D
        const int v3 = lane & (NW - 1);
        uint32_t w16[16];
#pragma unroll
        for (int s = 0; s < 3; s++) { tm_put(w16, s, pa_[s]); tm_put(w16, 3 + s, pg_[s]); }
        tm_put(w16, 6, inv4[0]); tm_put(w16, 7, inv4[1]);
        tm_st16(tb + TM_C, w16);
        tm_put(w16, 0, inv4[2]); tm_put(w16, 1, inv4[3]);
        // lane 31 forms A_w = rsep - kP Z30 with its own separator row; lane 0 forms B_w for the previous warp's separator
        tm_put(w16, 2, s_l3[wid][0]);
        tm_put(w16, 3, (wid > 0) ? s_l3[(wid + NW - 1) % NW][1] : 0.0);
        tm_put(w16, 4, (wid > 0) ? s_l3[(wid + NW - 1) % NW][2] : 0.0);
        tm_put(w16, 5, (wid > 0) ? s_l3[(wid + NW - 1) % NW][3] : 0.0);
        tm_put(w16, 6, s_minv[wid][v3]);
        tm_put(w16, 7, (wid > 0) ? s_minv[(wid + NW - 1) % NW][v3] : 0.0);
        tm_st16(tb + TM_C + 16, w16);
#pragma unroll
        for (int k = 0; k < 6; k++) tm_put(w16, k, gr[k + 1]);
        tm_put(w16, 6, GL); tm_put(w16, 7, GR);
        tm_st16(tb + TM_D + 16, w16);
        tm_wait_st();
      }
    }
    // ------------------------------------------------------------------ initial condition
#pragma unroll
    for (int k = 0; k < C; k++) { q[k] = (t * C + k < P.ni) ? 1.0 : 0.0; phi[k] = 0.0; }
    XL = (t > 0 && t * C - 1 < P.ni) ? 1.0 : 0.0;
    qn = ((t + 1) * C < P.ni) ? 1.0 : 0.0;
    const bool full = P.store_full != 0;
    double *hw = P.hist + (size_t)(full ? p : blockIdx.x) * P.hist_stride + 2 * t;   // write cursor (slice j)
    const double *hr = hw + (size_t)n * SL;                                        // read cursor (slice n-j)
    auto store_slice = [&](double *dst) {
#pragma unroll
      for (int k = 0; k < C; k += 2) *reinterpret_cast<double2 *>(dst + k * T) = make_double2(q[k], q[k + 1]);
    };
    store_slice(hw);
    const double *wq = P.w;
    constexpr unsigned QO_BUF = (unsigned)((C / 2) * T * 16);
    const unsigned qo_me = smem_u32(&s_qo[0][0][t][0]);
    const unsigned pub_mine = smem_u32(&s_pub[0][wid][0]);                       // lane 31: A_wid
    const unsigned pub_prev = smem_u32(&s_pub[0][(wid + NW - 1) % NW][1]);       // lane 0 of warp wid > 0: B_wid
    const unsigned pub_v = smem_u32(&s_pub[0][lane & (NW - 1)][0]);
    const int g8 = lane & 7;
    const bool pub0 = (lane == 0) && (wid > 0), pub31 = (lane == 31);

    uint32_t rA[16];
    tm_ld16(tb + TM_A, rA);
    tm_wait_ld();
    tm_pin(rA);

    // PH = 0: j < n/2 (store the slice), 1: j = n/2 (weight w_j, pairs with itself), 2: j > n/2 (pairs with slice n-j)
    auto step = [&](auto ph, const int j) {
      constexpr int PH = decltype(ph)::value;
      hw += SL; hr -= SL;
      if (PH >= 1) {   // the slices the next PD steps pair with are in flight; fetch the one for step j + PD
#pragma unroll
        for (int a = (PH == 1 ? 1 : PD); a <= PD; a++) {   // the middle step starts the pipeline: steps j+1 .. j+PD
          if (j + a <= n) {
            const unsigned dst = qo_me + ((j + a) & (RING - 1)) * QO_BUF;
            const double *src = hr - a * SL;
#pragma unroll
            for (int k = 0; k < C; k += 2) cp_async16(dst + (k / 2) * T * 16, src + k * T);
          }
          cp_async_commit();
        }
      }
      double wj = 0.0;
      if (PH >= 1) wj = __ldg(wq + j);
      uint32_t rB[32];
      tm_ld32(tb + TM_B, rB);
#if TM_TWOSIDED
      // ---- eliminations from both chunk ends towards node 3 (block A), on u = b / A_off
      double z[CI];
      {
        double tk[CI];
#pragma unroll
        for (int k = 0; k < CI; k++) tk[k] = fma(4.0, q[k], ((k == 0) ? XL : q[k - 1]) + q[k + 1]);
        const double y1 = fma(-tm_get(rA, 0), tk[0], tk[1]), y5 = fma(-tm_get(rA, 3), tk[6], tk[5]);
        const double y2 = fma(-tm_get(rA, 1), y1, tk[2]), y4 = fma(-tm_get(rA, 2), y5, tk[4]);
        const double y3 = fma(-tm_get(rA, 4), y2, fma(-tm_get(rA, 5), y4, tk[3]));
        z[3] = tm_get(rA, 6) * y3;
        z[0] = tk[0]; z[1] = y1; z[2] = y2; z[4] = y4; z[5] = y5; z[6] = tk[6];
      }
      const double sl = tm_get(rA, 7);
      // ---- substitution outwards from node 3 and separator row (block B)
      tm_wait_ld();
      tm_pin(rB);
      z[2] = fma(-tm_get(rB, 8), z[3], tm_get(rB, 2) * z[2]);
      z[4] = fma(-tm_get(rB, 9), z[3], tm_get(rB, 3) * z[4]);
      z[1] = fma(-tm_get(rB, 7), z[2], tm_get(rB, 1) * z[1]);
      z[5] = fma(-tm_get(rB, 10), z[4], tm_get(rB, 4) * z[5]);
      z[0] = fma(-tm_get(rB, 6), z[1], tm_get(rB, 0) * z[0]);
      z[6] = fma(-tm_get(rB, 11), z[5], tm_get(rB, 5) * z[6]);
      const double zfn = shfl_dn_d(z[0], 1);
#else
      // ---- first sweep (block A): u = b / A_off from the right end of the chunk
      double z[CI];
#pragma unroll
      for (int k = CI - 1; k >= 0; k--) {
        const double qm = (k == 0) ? XL : q[k - 1], qp = q[k + 1];
        const double tk = fma(4.0, q[k], qm + qp);
        z[k] = (k == CI - 1) ? tk : fma(-tm_get(rA, k), z[k + 1], tk);
      }
      z[0] = tm_get(rA, 6) * z[0];
      const double zfn = shfl_dn_d(z[0], 1);
      const double sl = tm_get(rA, 7);
      // ---- second sweep and separator row (block B)
      tm_wait_ld();
      tm_pin(rB);
#pragma unroll
      for (int k = 1; k < CI; k++) z[k] = fma(-tm_get(rB, 5 + k), z[k - 1], tm_get(rB, k - 1) * z[k]);
#endif
      const double Aoff = tm_get(rB, 12), su = tm_get(rB, 13);
      uint32_t rC[32];
      tm_ld32(tb + TM_C, rC);
      double r = Aoff * fma(4.0, q[C - 1], q[CI - 1]);
      const double rnx = fma(Aoff, qn, -su * zfn);
      r = fma(-sl, z[CI - 1], r);
      const double rsep = r;
      r = (lane == 31) ? 0.0 : r + rnx;
      // ---- level 2 (block C): three cyclic-reduction stages, then the 4x4 class inverse
      tm_wait_ld();
      tm_pin(rC);
#pragma unroll
      for (int s = 0; s < NST; s++) {
        const int d = 1 << s;
        const double rm = shfl_up_d(r, d), rp = shfl_dn_d(r, d);
        r = fma(tm_get(rC, s), rm, fma(tm_get(rC, 3 + s), rp, r));
      }
      double Z;
      {
        const double r0 = shfl_d(r, g8), r1 = shfl_d(r, g8 + 8), r2 = shfl_d(r, g8 + 16), r3 = shfl_d(r, g8 + 24);
        Z = fma(tm_get(rC, 6), r0, tm_get(rC, 7) * r1) + fma(tm_get(rC, 8), r2, tm_get(rC, 9) * r3);
      }
      // ---- level 3: lanes 31 / 0 publish the two halves of R_v, one barrier, lane-parallel 4x4 solve
      double Zm = shfl_up_d(Z, 1);
      const unsigned pbuf = (j & 1) * (unsigned)(NW * 2 * 8);
      {
        const double Apub = fma(-tm_get(rC, 10), Zm, rsep);
        const double Bpub = fma(tm_get(rC, 11), q[0], fma(-tm_get(rC, 12), z[0], -tm_get(rC, 13) * Z));
        if (pub31) sts64(pub_mine + pbuf, Apub);
        if (pub0) sts64(pub_prev + pbuf, Bpub);
      }
      Zm = (lane == 0) ? 0.0 : Zm;
      const double mW = tm_get(rC, 14), mM = tm_get(rC, 15);
      __syncthreads();
      uint32_t rD[32];
      tm_ld32(tb + TM_D, rD);
      tm_ld16(tb + TM_A, rA);       // next step's first-sweep coefficients
      double Ww, Wm;
      {
        const double2 ab = lds128(pub_v + pbuf);
        const double R = ab.x + ab.y;
        Ww = mW * R; Wm = mM * R;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
          Ww += __shfl_xor_sync(0xffffffffu, Ww, d);
          Wm += __shfl_xor_sync(0xffffffffu, Wm, d);
        }
      }
      tm_wait_ld();
      tm_pin(rD);
      tm_pin(rA);
      const double X = (lane == 31) ? Ww : fma(-tm_get(rD, 14), Wm, fma(-tm_get(rD, 15), Ww, Z));
      const double XLn = fma(-GLm, Wm, fma(-GRm, Ww, Zm));
#pragma unroll
      for (int k = 0; k < CI; k++) q[k] = fma(-tm_get(rD, k), XLn, fma(-tm_get(rD, 7 + k), X, z[k]));
      q[C - 1] = X;
      XL = XLn;
      qn = shfl_dn_d(q[0], 1);
      qn = (lane == 31) ? 0.0 : qn;
      // ---- history + fused quadrature
      if (PH == 0 || full) store_slice(hw);
      if (PH == 1) {
#pragma unroll
        for (int k = 0; k < C; k++) phi[k] = fma(wj * q[k], q[k], phi[k]);
      }
      if (PH == 2) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(PD) : "memory");   // all but the PD newest groups have landed
        const unsigned src = qo_me + (j & (RING - 1)) * QO_BUF;
#pragma unroll
        for (int k = 0; k < C; k += 2) {
          const double2 v = lds128(src + (k / 2) * T * 16);
          phi[k] = fma(wj * q[k], v.x, phi[k]);
          phi[k + 1] = fma(wj * q[k + 1], v.y, phi[k + 1]);
        }
      }
    };
    // two steps per loop trip: the second step's sweeps do not depend on the first step's history store / quadrature
    // update, so the scheduler can slide those under the dependent chains of the next step
    int j = 1;
    for (; 2 * (j + 1) < n; j += 2) { step(std::integral_constant<int, 0>{}, j); step(std::integral_constant<int, 0>{}, j + 1); }
    for (; 2 * j < n; j++) step(std::integral_constant<int, 0>{}, j);
    step(std::integral_constant<int, 1>{}, j);   // n is even: j = n/2
    for (j++; j + 1 <= n; j += 2) { step(std::integral_constant<int, 2>{}, j); step(std::integral_constant<int, 2>{}, j + 1); }
    for (; j <= n; j++) step(std::integral_constant<int, 2>{}, j);
    cp_async_wait0();

    // ------------------------------------------------------------------ residual, phi, Q
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
  if (wid == 0) tm_dealloc(s_tm);
}

}  // namespace scftb
