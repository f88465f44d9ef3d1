// march1d.cuh — the contour-march kernel of the 1D SCFT residual (sm_100a, fp64).
//
// One CTA per problem.  The whole evaluation
//     assemble (M + ds*K + ds*M_w)  ->  factor once  ->  n implicit-Euler steps  ->
//     phi(x) = int q(x,s) q(x,1-s) ds  ->  residual, Q
// runs inside one launch; q, the factor and the phi accumulators never leave the register file.
// It replaces, for the tridiagonal (1D / y-invariant strip) case, the reference's
//     assembly         1D_FEM.c:95-186, scft.cc:589-669
//     factorisation    KSPSetUp/PCICC 1D_FEM.c:191-206, UMFPACK scft.cc:695
//     contour loop     1D_FEM.c:208-228, drivescft.cc:130-165
//     quadrature       1D_FEM.c:260-277, drivescft.cc:184-213 (romint.c:21-57 as a weight vector)
//
// Parallel-in-x solve of the constant tridiagonal system (nested substructuring):
//   level 1  each thread owns C consecutive nodes: C-1 chunk-interior nodes + 1 separator.
//            Thomas on the chunk (pivots pre-inverted, rows pre-scaled), two precomputed spikes.
//   level 2  the 31 separators inside a warp: cyclic reduction in warp shuffles with
//            precomputed multipliers (5 stages), two precomputed warp spikes.
//   level 3  the warp separators (blockDim/32 unknowns): precomputed dense inverse, one
//            shared-memory exchange and ONE __syncthreads per contour step.
// All elimination coefficients are computed once per field update and reused for all n steps.
//
// History layout in HBM: slice j of a problem is T*C doubles, element (k,t) at k*T+t, so a warp
// stores/loads 256 contiguous bytes per instruction.  Only slices j < n/2 are re-read: Romberg
// weights are symmetric, so phi_i = sum_{j>n/2} 2 w_j q_i(j) q_i(n-j) + w_{n/2} q_i(n/2)^2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scftb {

struct MarchParams {
  int N;                  // nodes
  int ni;                 // interior nodes N-2
  int nsteps;             // contour steps n
  int scheme;             // 0 row-scaled C, 1 consistent C
  int nprob;              // problems in this launch
  int store_full;         // 1: store every slice, per problem; 0: half history, per CTA slot
  int uniform;            // 1: uniform mesh (x == nullptr)
  int pshare;             // 1: every problem of the launch uses parameter slot 0 (f0, L, x, chi) and writes only `out`
                          //    (the columns of a finite-difference Jacobian, fdjac.c:18-34); 0: problem p uses slot p
  double sign;
  const double *eta_mid;  // [nprob][ni], problem stride eta_stride
  long long eta_stride;   // doubles between consecutive problems in eta_mid
  long long out_stride;   // same for out
  const int *skip;        // [nprob] non-zero: leave the problem untouched (converged) or nullptr
  const double *f0;       // [nprob][N]
  const double *L;        // [nprob]
  const double *x;        // [nprob][N] node coordinates (non-uniform) or nullptr
  const double *eta_bnd;  // [nprob][2] spline-extrapolated wall values (non-uniform) or nullptr
  const double *w;        // [nsteps+1] quadrature weights
  double *hist;           // history slices
  long long hist_stride;  // doubles per problem (store_full) or per CTA slot
  double *out;            // [nprob][ni]  sign*(phi0 - phi)
  double *phi;            // [nprob][N]
  double *Q;              // [nprob]
  double *eta_full;       // [nprob][N] (diagnostic, scft.cc:452-490) or nullptr
  // two-species instantiation only (TWO = true): fields [nprob][2][ni] in eta_mid (eta_stride >= 2 ni)
  int jf;                 // contour steps of the A block, 0 < jf < nsteps
  const double *wA, *wB;  // [nsteps+1] block quadrature weights over the contour index (zero outside the block)
  const double *chi;      // [nprob] chi N
  const double *eta_bndB; // [nprob][2] wall values of the B field (non-uniform) or nullptr
  double *phiB;           // [nprob][N]
};

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_dn_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_d(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// shared-memory accesses through precomputed 32-bit shared-window addresses: keeps the per-step
// address arithmetic out of the contour loop
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts128(unsigned a, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double2 lds128(unsigned a) {
  double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory"); return v;
}

// Ampere-style asynchronous global->shared copies (LDGSTS)
__device__ __forceinline__ void cp_async16(unsigned sa, const void *gmem) {
  // .cg (L1 bypass): the .ca flavour measured 2.6 % slower on the 4096-problem sweep (16.59 vs 16.16 ms)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned sa, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// eta on node i (0..N-1).  Wall nodes: natural-spline extrapolation of the interior values
// (scft.cc:456-475 -> spline_chen.c:77-100).  With y''=0 at the end knots the cubic term of the
// first/last piece vanishes on a uniform mesh, leaving the linear extrapolation a*y0+b*y1.
__device__ __forceinline__ double eta_node(const MarchParams &P, int p, int i, double L, bool fieldB = false) {
  const double *em = P.eta_mid + (size_t)p * P.eta_stride + (fieldB ? P.ni : 0);
  if (i >= 1 && i <= P.N - 2) return em[i - 1];
  if (!P.uniform) return (fieldB ? P.eta_bndB : P.eta_bnd)[2 * p + (i == 0 ? 0 : 1)];
  const int N = P.N;
  if (i == 0) {
    double x0 = L * 0 / (N - 1), x1 = L * 1 / (N - 1), x2 = L * 2 / (N - 1);
    double h = x2 - x1, a = (x2 - x0) / h, b = (x0 - x1) / h;
    return a * em[0] + b * em[1];
  }
  double xe = L * (N - 1) / (N - 1), xa = L * (N - 3) / (N - 1), xb = L * (N - 2) / (N - 1);
  double h = xb - xa, a = (xb - xe) / h, b = (xe - xa) / h;
  return a * em[P.ni - 2] + b * em[P.ni - 1];
}

struct Row { double Al, Ad, Au, Tl, Td, Tu, Dl, Dd, Du; };   // A (mass), T = A + dt*D, D = B + C

// Row g (0-based interior index) of A (mass) and T = A + ds*(B + C).
__device__ __forceinline__ Row assemble_row(const MarchParams &P, int p, int g, double L, double dt, bool fieldB = false) {
  Row r;
  if (g >= P.ni) { r.Al = r.Ad = r.Au = 0.0; r.Tl = r.Tu = 0.0; r.Td = 1.0; r.Dl = r.Du = 0.0; r.Dd = 0.0; return r; }  // padding
  const int i = g + 1;
  double a1, a2, bl, bd, bu;
  if (P.uniform) {  // 1D_FEM.c:61,95-98
    double h = L / (P.N - 1);
    a1 = a2 = h;
    r.Al = h / 6; r.Ad = 2. / 3 * h; r.Au = h / 6;
    bl = -1 / h; bd = 2. / h; bu = -1 / h;
  } else {          // simple_FEM_1D_transient.m:35-57
    const double *x = P.x + (size_t)(P.pshare ? 0 : p) * P.N;
    a1 = x[i] - x[i - 1]; a2 = x[i + 1] - x[i];
    r.Al = a1 / 6; r.Ad = a1 / 3 + a2 / 3; r.Au = a2 / 6;
    bl = -1 / a1; bd = 1 / a1 + 1 / a2; bu = -1 / a2;
  }
  double e0 = eta_node(P, p, i, L, fieldB), cl, cd, cu;
  if (P.scheme == 0) {  // row-scaled lumping, 1D_FEM.c:104-105
    cl = r.Al * e0; cd = r.Ad * e0; cu = r.Au * e0;
  } else {              // (eta_h phi_i, phi_j), 2-point Gauss exact for linear eta_h (scft.cc:653-655)
    double em = eta_node(P, p, i - 1, L, fieldB), ep = eta_node(P, p, i + 1, L, fieldB);
    cl = a1 * (em + e0) / 12;
    cu = a2 * (e0 + ep) / 12;
    cd = a1 * (em + 3 * e0) / 12 + a2 * (3 * e0 + ep) / 12;
  }
  r.Dl = bl + cl; r.Dd = bd + cd; r.Du = bu + cu;
  r.Tl = r.Al + dt * r.Dl;
  r.Td = r.Ad + dt * r.Dd;
  r.Tu = r.Au + dt * r.Du;
  if (g == 0) { r.Al = 0.0; r.Tl = 0.0; r.Dl = 0.0; }            // q(0,s) = 0   (1D_FEM.c:179-180,213)
  if (g == P.ni - 1) { r.Au = 0.0; r.Tu = 0.0; r.Du = 0.0; }     // q(L,s) = 0   (1D_FEM.c:181-182,214)
  return r;
}

// index of node k (0..C-1) of thread t inside a history slice of T*C doubles: pairs of
// consecutive nodes of a thread are adjacent so that a warp moves 512 contiguous bytes per
// 128-bit instruction (C = 1: 256 bytes per 64-bit instruction).
__host__ __device__ inline int hist_index(int C, int T, int t, int k) {
  return (C == 1) ? t : ((k >> 1) * T + t) * 2 + (k & 1);
}

constexpr int PUB = 10; // doubles per warp in a publish buffer: [0] qf, [1] zf, [2] Z0 (lane 0); [4] Z30 (lane 30), [5] rsep (lane 31);
                        // 80-byte rows: the four rows a warp reads in level 3 fall into disjoint banks (64-byte rows conflict 2-way)

// ODDN: instantiation for an odd number of contour steps (only then the first pairing step needs its partner slice
// re-read; keeping it out of the even-n instantiation leaves the hot loop's register allocation untouched)
// TWO: two-species (AB diblock) instantiation, SURVEY.md section 8(f)-4 — not in the reference.  The march runs as four
// segments, each with its own assembly + factor setup: q forward through the A block (field A, steps 1..jf) and the B block
// (field B), every slice stored; then q+ backward through the B block and the A block, every step paired with the stored
// slice q(n-j) and accumulated into phi_A / phi_B with the block weights.  TWO = false is the reference's one-sweep form.
template <int C, int T, bool UNI, int MINB, bool ODDN, bool TWO = false>
__global__ void __launch_bounds__(T, MINB) march_ie_kernel(MarchParams P) {
  static_assert(C == 1 || (C % 2) == 0, "C must be 1 or even");
  constexpr int CI = C - 1;             // chunk-interior nodes per thread
  constexpr int CA = CI > 0 ? CI : 1;   // array extent
  constexpr int NW = T / 32;
  constexpr int SL = T * C;             // doubles per history slice
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  __shared__ __align__(16) double s_ex[2][T];        // setup exchange gl0 / gr0, then scratch
  __shared__ __align__(16) double s_l3[NW][4];       // per warp separator v: P, cAu, csu, Nx
  __shared__ __align__(16) double s_l3s[NW][6];      // setup only: D, GL0, GR0, GL30, GR30
  __shared__ __align__(16) double s_minv[NW][NW];    // inverse of the level-3 matrix
  __shared__ __align__(16) double s_pub[2][NW][PUB]; // per-step publish buffers (double-buffered)
  __shared__ double s_red[NW];
  // staging ring for the paired history slice q(., n-j): each thread cp.async's and reads back its
  // own 16-byte pieces, so no barrier is involved
  constexpr bool STAGE = (C * T <= 2048);   // 32 KB of static shared memory at most
  __shared__ __align__(16) double s_qo[STAGE ? 2 : 1][STAGE ? (C + 1) / 2 : 1][STAGE ? T : 1][2];
  const int n = P.nsteps;
  const double dt = 1.0 / n;            // time_step = 1/(total_time_step-1), scft.cc:29

  for (int p = blockIdx.x; p < P.nprob; p += gridDim.x) {
    if (P.skip && P.skip[p]) continue;
    const int pp = P.pshare ? 0 : p;      // parameter slot
    const double L = P.L[pp];
    // state that lives across the segments of a two-species march (one segment otherwise)
    double q[C], phi[C], phiB[TWO ? C : 1];
    double XL, qn, qsumF = 0.0;
    double *hw;
    const double *hr;
    for (int seg = 0; seg < (TWO ? 4 : 1); seg++) {
    const bool fB = TWO && (seg == 1 || seg == 2);        // field of this segment
    const bool backward = TWO && seg >= 2;                // q+ sweep
    // ------------------------------------------------------------------ assembly + level 1
    double ca[CA], cd[CA], cu[CA];      // pre-scaled rows of A on chunk-interior nodes (UNI: ca==cu)
    double al[CA], be[CA], gl[CA], gr[CA];
    double sAl, sAd, sAu, sl, sd, su;   // separator row of A and T
    {
      Row rs = assemble_row(P, p, t * C + CI, L, dt, fB);
      sAl = rs.Al; sAd = rs.Ad; sAu = rs.Au; sl = rs.Tl; sd = rs.Td; su = rs.Tu;
      // UNI: A's separator row is A_off * (1, 4, 1); the Dirichlet zeroing is carried by the neighbour
      // values (exactly 0 on walls / padding), so one coefficient is kept (in sAd)
      if (UNI) sAd = (rs.Al != 0.0) ? rs.Al : rs.Au;
    }
    if constexpr (CI > 0) {
      // UL order: eliminate from the chunk's right end towards the left, then substitute left to
      // right.  z_first is then known after the FIRST sweep (its shuffle to the left neighbour
      // overlaps the second sweep) and z_last, which the own separator needs, comes out last.
      double Tl0 = 0, TuL = 0, pinv_next = 0, Tl_next = 0;
      double ulo[CA];                           // pinv_k * Tu_k (setup only)
#pragma unroll
      for (int k = CI - 1; k >= 0; k--) {
        Row r = assemble_row(P, p, t * C + k, L, dt, fB);
        double piv = (k == CI - 1) ? r.Td : r.Td - (r.Tu * pinv_next) * Tl_next;
        double pinv = 1.0 / piv;
        // UNI: one off-diagonal coefficient serves both neighbours; the Dirichlet zeroing of A is
        // then carried by the neighbour values themselves (wall / padding nodes are exactly 0),
        // and A's diagonal is 4x the off-diagonal, so no cd[] is kept.
        ca[k] = pinv * ((UNI && r.Al == 0.0) ? r.Au : r.Al); cd[k] = pinv * r.Ad; cu[k] = pinv * r.Au;
        ulo[k] = (k == CI - 1) ? 0.0 : pinv * r.Tu;
        // first sweep; UNI runs it on u = y/A_off with the plain multiplier Tu_k/piv_{k+1}
        al[k] = (k == CI - 1) ? 0.0 : (UNI ? r.Tu * pinv_next : ulo[k]);
        be[k] = (k == 0) ? 0.0 : pinv * r.Tl;   // second sweep: coupling to node k-1
        if (k == 0) Tl0 = pinv * r.Tl;          // scaled coupling to the left separator
        if (k == CI - 1) TuL = pinv * r.Tu;     // scaled coupling to the own (right) separator
        pinv_next = pinv; Tl_next = r.Tl;
      }
      // spikes gl = T_loc^-1 (Tl_first e_first), gr = T_loc^-1 (Tu_last e_last)
      gl[0] = Tl0;
#pragma unroll
      for (int k = 1; k < CI; k++) gl[k] = -be[k] * gl[k - 1];
      double y[CA];
      y[CI - 1] = TuL;
#pragma unroll
      for (int k = CI - 2; k >= 0; k--) y[k] = -ulo[k] * y[k + 1];
      gr[0] = y[0];
#pragma unroll
      for (int k = 1; k < CI; k++) gr[k] = y[k] - be[k] * gr[k - 1];
    }
    // ------------------------------------------------------------------ Schur rows on separators
    double a, b, c;
    __syncthreads();  // previous problem's readers of shared memory are done
    if constexpr (CI > 0) {
      s_ex[0][t] = gl[0]; s_ex[1][t] = gr[0];
      __syncthreads();
      double gl0n = (t + 1 < T) ? s_ex[0][t + 1] : 0.0, gr0n = (t + 1 < T) ? s_ex[1][t + 1] : 0.0;
      a = -sl * gl[CI - 1];
      b = sd - sl * gr[CI - 1] - su * gl0n;
      c = -su * gr0n;
    } else {
      a = sl; b = sd; c = su;
    }
    // ------------------------------------------------------------------ level 2: warp PCR setup
    const double l3P = a, l3D = b, l3N = c;       // lane 31: row of the warp separator
    const double A0 = (lane == 0) ? a : 0.0, C30 = (lane == 30) ? c : 0.0;
    if (lane == 31) { a = 0.0; b = 1.0; c = 0.0; }
    if (lane == 0) a = 0.0;
    if (lane == 30) c = 0.0;
    // three cyclic-reduction stages (strides 1, 2, 4) leave eight independent 4-unknown systems, one
    // per residue class lane mod 8; each lane keeps its row of that system's inverse
    constexpr int NST = 3;
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
      // rows of the class system: lanes g, g+8, g+16, g+24; row m of M^-1 is the solution of M^T x = e_m
      const int g8 = lane & 7, m8 = lane >> 3;
      double ra[4], rb[4], rc[4];
#pragma unroll
      for (int mm = 0; mm < 4; mm++) { ra[mm] = shfl_d(a, g8 + 8 * mm); rb[mm] = shfl_d(b, g8 + 8 * mm); rc[mm] = shfl_d(c, g8 + 8 * mm); }
      // M^T is tridiagonal with sub-diagonal rc[i-1] (row i, col i-1), diagonal rb[i], super-diagonal ra[i+1]
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
    // left neighbour's spikes, so that its separator value can be formed locally (no shuffle on the
    // critical path); lane 0's left neighbour is the previous warp's separator W_{wid-1} itself
    double GLm = shfl_up_d(GL, 1), GRm = shfl_up_d(GR, 1);
    if (lane == 0) { GLm = -1.0; GRm = 0.0; }
    // ------------------------------------------------------------------ level 3 setup
    if (lane == 31) {
      s_l3[wid][0] = l3P; s_l3[wid][1] = (wid + 1 < NW) ? (UNI ? sAd : sAu) : 0.0;
      s_l3[wid][2] = (CI > 0) ? su : 0.0; s_l3[wid][3] = l3N;
      s_l3s[wid][0] = l3D;
    }
    if (lane == 0) { s_l3s[wid][1] = GL; s_l3s[wid][2] = GR; }
    if (lane == 30) { s_l3s[wid][3] = GL; s_l3s[wid][4] = GR; }
    __syncthreads();
    if (t < NW) {  // thread v: column v of M^-1 by Thomas (M is tridiagonal, diagonally dominant)
      double cc[NW], dd[NW];
      double cprev = 0.0, dprev = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        double Pw = s_l3[w][0], Dw = s_l3s[w][0], Nw = s_l3[w][3];
        double lo = (w > 0) ? -Pw * s_l3s[w][3] : 0.0;                                      // M[w][w-1]
        double di = Dw - Pw * s_l3s[w][4] - ((w + 1 < NW) ? Nw * s_l3s[(w + 1) % NW][1] : 0.0);
        double up = (w + 1 < NW) ? -Nw * s_l3s[(w + 1) % NW][2] : 0.0;                      // M[w][w+1]
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
    // rows wid and wid-1 of M^-1 are warp-uniform; keep them out of the loop's shared-memory traffic
    // only when they are few
    // ------------------------------------------------------------------ initial condition
    // q = 1 on interior nodes, 0 on walls / padding (drivescft.cc:120-127, 1D_FEM.c:114-129)
    const bool sweep_start = !TWO || seg == 0 || seg == 2;
    if (sweep_start) {
#pragma unroll
      for (int k = 0; k < C; k++) q[k] = (t * C + k < P.ni) ? 1.0 : 0.0;
      XL = (t > 0 && t * C - 1 < P.ni) ? 1.0 : 0.0;   // value at the left separator
      qn = ((t + 1) * C < P.ni) ? 1.0 : 0.0;          // first node of the next chunk
      double *hb = P.hist + (size_t)(P.store_full ? p : blockIdx.x) * P.hist_stride;
      // per-thread view of a slice: C == 1: doubles at [t]; else double2 at [(k/2)*T + t]
      hw = hb + ((C == 1) ? t : 2 * t);               // write cursor (slice j)
      hr = hw + (size_t)n * SL;                       // read cursor (slice n-j)
    }
    if (!TWO || seg == 0) {
#pragma unroll
      for (int k = 0; k < C; k++) { phi[k] = 0.0; if (TWO) phiB[k] = 0.0; }
    }
    auto store_slice = [&](double *dst) {
      if constexpr (C == 1) dst[0] = q[0];
      else {
#pragma unroll
        for (int k = 0; k < C; k += 2) *reinterpret_cast<double2 *>(dst + k * T) = make_double2(q[k], q[k + 1]);
      }
    };
    if (!TWO || seg == 0) store_slice(hw);
    const double *wq = P.w;
    const bool full = P.store_full != 0;

    // shared-window addresses and the lane's share of the level-3 solve (loop invariants)
    constexpr unsigned QO_BUF = (unsigned)(((C + 1) / 2) * T * 16);   // bytes per staging buffer
    constexpr unsigned PUB_BUF = (unsigned)(NW * PUB * 8);
    const unsigned qo_me = smem_u32(&s_qo[0][0][STAGE ? t : 0][0]);
    const unsigned pub_me = smem_u32(&s_pub[0][wid][0]);
    const int v3 = lane & (NW - 1);                // this lane evaluates separator row v3 of level 3
    const unsigned pub_v = smem_u32(&s_pub[0][v3][0]), pub_vn = smem_u32(&s_pub[0][(v3 + 1) % NW][0]);
    const double c3P = s_l3[v3][0], c3Au = s_l3[v3][1], c3su = s_l3[v3][2], c3Nx = s_l3[v3][3];
    const double mW = s_minv[wid][v3], mM = (wid > 0) ? s_minv[(wid + NW - 1) % NW][v3] : 0.0;
    const int g8 = lane & 7;
    auto prefetch = [&](int jj, const double *src) {   // slice n-jj -> staging buffer jj&1
      if (STAGE && (TWO ? backward : 2 * jj > n) && jj <= n) {
        const unsigned dst = qo_me + (jj & 1) * QO_BUF;
        if constexpr (C == 1) cp_async8(dst, src);
        else {
#pragma unroll
          for (int k = 0; k < C; k += 2) cp_async16(dst + (k / 2) * T * 16, src + k * T);
        }
      }
      if (STAGE) cp_async_commit();
    };
    if (TWO && seg == 2) {   // q+(., 0) = 1 pairs with the last slice of q (plain loads: stored by this thread)
      const double wa = __ldg(P.wA + n), wb = __ldg(P.wB + n);
#pragma unroll
      for (int k = 0; k < C; k++) {
        const double v = hr[(C == 1) ? 0 : (k / 2) * 2 * T + (k & 1)] * q[k];
        phi[k] = fma(wa, v, phi[k]);
        phiB[TWO ? k : 0] = fma(wb, v, phiB[TWO ? k : 0]);
      }
    }
    if (sweep_start) prefetch(1, hr - SL);
    // contour steps of this segment
    const int jbeg = !TWO ? 1 : (seg == 0 ? 1 : (seg == 1 ? P.jf + 1 : (seg == 2 ? 1 : n - P.jf + 1)));
    const int jend = !TWO ? n : (seg == 0 ? P.jf : (seg == 1 ? n : (seg == 2 ? n - P.jf : n)));

    // ------------------------------------------------------------------ the contour march
    for (int j = jbeg; j <= jend; j++) {
      hw += SL; hr -= SL;
      const bool pairing = TWO ? backward : (2 * j > n);
      // the slice the NEXT step pairs with is fetched a whole step ahead.  For odd n the first pairing step
      // (j = (n+1)/2) pairs with the slice stored one step earlier, which was not written yet when its prefetch
      // was issued: once, copy it into the staging buffer with ordinary loads (ordered after this thread's store)
      if (ODDN && STAGE && 2 * j == n + 1) {
        cp_async_wait0();   // the stale prefetch into this buffer has landed and can be overwritten
        const unsigned dst = qo_me + (j & 1) * QO_BUF;
        if constexpr (C == 1) sts64(dst, hr[0]);
        else {
#pragma unroll
          for (int k = 0; k < C; k += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(hr + k * T);
            sts128(dst + (k / 2) * T * 16, v.x, v.y);
          }
        }
      }
      prefetch(j + 1, hr - SL);
      // right-hand side b = A q and the level-1 (chunk) solve with zero separators, UL order
      double z[CA];
      double zlast = 0.0, z0 = 0.0, zfn = 0.0;
      if constexpr (CI > 0) {
        if constexpr (UNI) {
          // b_k = A_off (q_{k-1} + q_{k+1} + 4 q_k); first sweep on u = b/A_off from the right end
#pragma unroll
          for (int k = CI - 1; k >= 0; k--) {
            double qm = (k == 0) ? XL : q[k - 1], qp = q[k + 1];
            double tk = fma(4.0, q[k], qm + qp);
            z[k] = (k == CI - 1) ? tk : fma(-al[k], z[k + 1], tk);
          }
          z[0] = ca[0] * z[0];
          zfn = shfl_dn_d(z[0], 1);            // z_first of the right neighbour: overlaps the second sweep
#pragma unroll
          for (int k = 1; k < CI; k++) z[k] = fma(-be[k], z[k - 1], ca[k] * z[k]);
        } else {
#pragma unroll
          for (int k = CI - 1; k >= 0; k--) {
            double qm = (k == 0) ? XL : q[k - 1], qp = q[k + 1];
            double bk = fma(ca[k], qm, fma(cu[k], qp, cd[k] * q[k]));
            z[k] = (k == CI - 1) ? bk : fma(-al[k], z[k + 1], bk);
          }
          zfn = shfl_dn_d(z[0], 1);
#pragma unroll
          for (int k = 1; k < CI; k++) z[k] = fma(-be[k], z[k - 1], z[k]);
        }
        zlast = z[CI - 1]; z0 = z[0];
      }
      const double qprev = (CI > 0) ? q[CI > 0 ? CI - 1 : 0] : XL;
      double r = UNI ? sAd * fma(4.0, q[C - 1], qprev) : fma(sAl, qprev, sAd * q[C - 1]);   // UNI: sAd holds A_off
      const double rnx = fma(UNI ? sAd : sAu, qn, (CI > 0) ? -su * zfn : 0.0);             // next-chunk terms
      if constexpr (CI > 0) r = fma(-sl, zlast, r);
      const double rsep = r;                              // lane 31: without next-warp terms
      r = (lane == 31) ? 0.0 : r + rnx;
      // level 2: three cyclic-reduction stages, then the 4x4 class inverse
#pragma unroll
      for (int s = 0; s < NST; s++) {
        const int d = 1 << s;
        double rm = shfl_up_d(r, d), rp = shfl_dn_d(r, d);
        r = fma(pa_[s], rm, fma(pg_[s], rp, r));
      }
      double Z;
      {
        double r0 = shfl_d(r, g8), r1 = shfl_d(r, g8 + 8), r2 = shfl_d(r, g8 + 16), r3 = shfl_d(r, g8 + 24);
        Z = fma(inv4[0], r0, inv4[1] * r1) + fma(inv4[2], r2, inv4[3] * r3);
      }
      // level 3: publish, one barrier, then every group of NW lanes solves the warp-separator system
      // cooperatively (lane v3 forms R_v3, butterfly sum over the group) -> every lane holds W_wid, W_wid-1
      const unsigned pbuf = (j & 1) * PUB_BUF;
      asm volatile(
          "{\n .reg .pred p0, p30, p31;\n"
          " setp.eq.s32 p0, %0, 0;\n setp.eq.s32 p30, %0, 30;\n setp.eq.s32 p31, %0, 31;\n"
          " @p0 st.shared.v2.f64 [%1], {%2, %3};\n @p0 st.shared.f64 [%1+16], %4;\n"
          " @p30 st.shared.f64 [%1+32], %4;\n @p31 st.shared.f64 [%1+40], %5;\n}"
          ::"r"(lane), "r"(pub_me + pbuf), "d"(q[0]), "d"(z0), "d"(Z), "d"(rsep) : "memory");
      double Zm = shfl_up_d(Z, 1);                 // left neighbour's warp-local separator value
      Zm = (lane == 0) ? 0.0 : Zm;
      if constexpr (NW > 1) __syncthreads(); else __syncwarp();
      double Ww, Wm;
      {
        const double2 own = lds128(pub_v + pbuf + 32);     // Z30_v, rsep_v
        const double2 nxt = lds128(pub_vn + pbuf);         // qf_{v+1}, zf_{v+1}
        const double z0n = lds64(pub_vn + pbuf + 16);      // Z0_{v+1}
        double R = fma(-c3P, own.x, own.y);
        double R2 = fma(c3Au, nxt.x, -c3su * nxt.y);
        R2 = fma(-c3Nx, z0n, R2);
        R += R2;
        Ww = mW * R; Wm = mM * R;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
          Ww += __shfl_xor_sync(0xffffffffu, Ww, d);
          Wm += __shfl_xor_sync(0xffffffffu, Wm, d);
        }
      }
      const double X = (lane == 31) ? Ww : fma(-GL, Wm, fma(-GR, Ww, Z));
      const double XLn = fma(-GLm, Wm, fma(-GRm, Ww, Zm));   // lane 0: GLm = -1, GRm = Zm = 0 -> W_{wid-1}
      // level-1 correction
      if constexpr (CI > 0) {
#pragma unroll
        for (int k = 0; k < CI; k++) q[k] = fma(-gl[k], XLn, fma(-gr[k], X, z[k]));
      }
      q[C - 1] = X;
      XL = XLn;
      qn = shfl_dn_d(q[0], 1);
      qn = (lane == 31) ? 0.0 : qn;   // supplied through the publish buffer
      // history + fused quadrature
      if constexpr (TWO) {
        if (!backward) store_slice(hw);
        else {
          const double wa = __ldg(P.wA + (n - j)), wb = __ldg(P.wB + (n - j));   // weights of the q slice's contour index
          if (STAGE) cp_async_wait1();
          const unsigned src = qo_me + (j & 1) * QO_BUF;
          if constexpr (C == 1) {
            const double v = (STAGE ? lds64(src) : hr[0]) * q[0];
            phi[0] = fma(wa, v, phi[0]); phiB[0] = fma(wb, v, phiB[0]);
          } else {
#pragma unroll
            for (int k = 0; k < C; k += 2) {
              const double2 v = STAGE ? lds128(src + (k / 2) * T * 16) : *reinterpret_cast<const double2 *>(hr + k * T);
              const double v0 = v.x * q[k], v1 = v.y * q[k + 1];
              phi[k] = fma(wa, v0, phi[k]); phiB[TWO ? k : 0] = fma(wb, v0, phiB[TWO ? k : 0]);
              phi[k + 1] = fma(wa, v1, phi[k + 1]); phiB[TWO ? k + 1 : 0] = fma(wb, v1, phiB[TWO ? k + 1 : 0]);
            }
          }
        }
      } else {
      if (full || 2 * j < n) store_slice(hw);
      if (2 * j >= n) {
        const double wj = __ldg(wq + j);   // j > n/2: 2*w_j (pair j, n-j); j == n/2: w_j
        if (pairing) {
          if (STAGE) cp_async_wait1();   // everything but the newest group (step j+1) has landed
          const unsigned src = qo_me + (j & 1) * QO_BUF;
          if constexpr (C == 1) phi[0] = fma(wj * q[0], STAGE ? lds64(src) : hr[0], phi[0]);
          else {
#pragma unroll
            for (int k = 0; k < C; k += 2) {
              const double2 v = STAGE ? lds128(src + (k / 2) * T * 16) : *reinterpret_cast<const double2 *>(hr + k * T);
              phi[k] = fma(wj * q[k], v.x, phi[k]);
              phi[k + 1] = fma(wj * q[k + 1], v.y, phi[k + 1]);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < C; k++) phi[k] = fma(wj * q[k], q[k], phi[k]);
        }
      }
      }
    }
    if (TWO && seg == 1) {   // Q = (1/L) int q(x,1) dx from the forward sweep
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int g = t * C + k;
        if (g < P.ni) {
          const int i = g + 1;
          double hw2;
          if (P.uniform) { double h = L / (P.N - 1); hw2 = 0.5 * (h + h); }
          else { const double *x = P.x + (size_t)pp * P.N; hw2 = 0.5 * ((x[i] - x[i - 1]) + (x[i + 1] - x[i])); }
          qsumF += hw2 * q[k];
        }
      }
    }
    }   // segments

    // ------------------------------------------------------------------ residual, phi, Q
    double qsum = 0.0;
    if constexpr (TWO) {
      const double chi = P.chi[pp];
      const double *em = P.eta_mid + (size_t)p * P.eta_stride;
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int g = t * C + k;
        if (g < P.ni) {
          const int i = g + 1;
          const double f0 = P.f0[(size_t)pp * P.N + i], pa = phi[k], pb = phiB[TWO ? k : 0];
          P.out[(size_t)p * P.out_stride + g] = P.sign * (f0 - pa - pb);
          P.out[(size_t)p * P.out_stride + P.ni + g] = em[g] - em[P.ni + g] - chi * (pb - pa);
          if (!P.pshare) {
            P.phi[(size_t)p * P.N + i] = pa;
            P.phiB[(size_t)p * P.N + i] = pb;
          }
        }
      }
      if (t == 0 && !P.pshare) {
        P.phi[(size_t)p * P.N] = 0.0; P.phi[(size_t)p * P.N + P.N - 1] = 0.0;
        P.phiB[(size_t)p * P.N] = 0.0; P.phiB[(size_t)p * P.N + P.N - 1] = 0.0;
      }
      qsum = qsumF;
    } else {
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int g = t * C + k;
      if (g < P.ni) {
        const int i = g + 1;
        const double f0 = P.f0[(size_t)pp * P.N + i];
        P.out[(size_t)p * P.out_stride + g] = P.sign * (f0 - phi[k]);
        if (!P.pshare) P.phi[(size_t)p * P.N + i] = phi[k];
        double hw2;
        if (P.uniform) { double h = L / (P.N - 1); hw2 = 0.5 * (h + h); }
        else { const double *x = P.x + (size_t)pp * P.N; hw2 = 0.5 * ((x[i] - x[i - 1]) + (x[i + 1] - x[i])); }
        qsum += hw2 * q[k];
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
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, d);
    if (lane == 0) s_red[wid] = qsum;
    __syncthreads();
    if (t == 0 && !P.pshare) {
      double s = 0.0;
      for (int w = 0; w < NW; w++) s += s_red[w];
      double len = P.uniform ? L : (P.x[(size_t)pp * P.N + P.N - 1] - P.x[(size_t)pp * P.N]);
      P.Q[p] = s / len;
    }
  }
}

}  // namespace scftb
