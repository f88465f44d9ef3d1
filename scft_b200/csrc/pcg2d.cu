// pcg2d.cu — the 2-D path: Q1 finite elements on a structured nx x ny mesh of [0,L] x [0,Ly],
// implicit-Euler contour march with a Jacobi-preconditioned conjugate-gradient solve per step
// (BASELINE.json configs[3], [4]).
//
// Reference: the deal.II matrices A, B, C of scft.cc:643-656 (2x2 Gauss on Q1 cells, Dirichlet on
// x = 0, L via scft.cc:599-606), sparsity from setup_system (scft.cc:546-549), and the CG-per-step
// pattern of the dead solve_time_step (scft.cc:698-705) / step-26.cc:224-241.  The live reference
// stepper (IRK4 block system) is non-symmetric, so the CG-able form is the implicit-Euler step
// (A + ds (B + C)) q+ = A q, which is SPD.
//
// Design
//   * assembly: one thread per row writes the 9-point rows of T = A + ds(B+C) and A straight into a
//     sliced-ELL layout (slices of 32 rows, 9 slots, slot-major inside a slice) so that a warp's
//     loads of values and column indices are contiguous; the CSR view is exported for inspection.
//   * solve: Chronopoulos-Gear CG (one fused reduction per iteration) with a Jacobi preconditioner.
//     Single GPU: the WHOLE contour march — right-hand sides, all CG iterations of all steps,
//     history stores and the fused density quadrature — is ONE persistent cooperative kernel with
//     grid-wide barriers; reductions are per-block partials summed in a fixed order by every block
//     (deterministic, no atomics).
//   * multi GPU: slab partition in x (contiguous row blocks, one node column of halo per side);
//     the same device code runs as two short kernels per CG iteration with an NCCL halo exchange
//     (grouped ncclSend/ncclRecv to the <= 2 neighbours) and ONE ncclAllReduce of 3 doubles between
//     them.  Convergence is decided on the device from the all-reduced scalars, identically on
//     every rank; the host polls the flag every few iterations.
#include <cooperative_groups.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine.h"

namespace cg = cooperative_groups;
using namespace scftb;

namespace scftb {

constexpr int SLOTS = 9;
constexpr int TPB2 = 256;
#ifndef MINB2D
#define MINB2D 4   // resident blocks per SM the single-GPU persistent kernel is compiled for (register cap 64)
#endif

struct Sys2D {
  int nx, ny, nyp;          // cells, nodes per column
  int row0, nrows;          // first owned global row, owned rows (whole node columns)
  int halo;                 // halo entries on each side (nyp or 0)
  int nslices;              // ceil(nrows/32)
  double hx, hy, dt;
  const double *eta;        // [ndof_global] field on all nodes
  int *col;                 // SELL: [nslices][9][32], LOCAL vector index (halo_left + owned + halo_right)
  double *valT, *valA;      // same layout
  double *dinv;             // [nrows] 1 / diag(T)
  // matrix-free application (apply_row): element entries of A and of A + dt B by the relation of the two nodes inside a
  // cell (0 same node, 1 x-neighbour, 2 y-neighbour, 3 diagonal), and dt hx hy / 144 for the (eta_h phi_a, phi_b) part
  double eA[4], eT[4], cC;
};

struct Vec2D {
  double *q, *x, *r, *z, *s, *p, *w, *b;   // z carries halos: z[-halo .. nrows+halo); others [nrows]
  double *qp, *w1;                         // q_{j-2}, T q_{j-2}: the extrapolated initial guess of a step (step_guess)
  double *qpp, *w2;                        // q_{j-3}, T q_{j-3} (quadratic extrapolation; nullptr = linear only)
};

struct Red2D {
  double *partial;          // [2][grid][4] per-block partial sums
  double *scal;             // [0..3] reduced gamma, delta, rr, bb; then two state records of 8 doubles at [8], [16]:
                            //   0 gamma_old, 1 alpha_old, 2 converged, 3 bb, 4 CG iterations done in this step
};

__device__ __forceinline__ size_t sell(int row, int k) { return (size_t)(row >> 5) * (SLOTS * 32) + k * 32 + (row & 31); }

// ------------------------------------------------------------------------------------------------
// assembly of row `i` (local) = global row row0 + i
__global__ void assemble2d_kernel(Sys2D S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nslices * 32) return;
  double vT[SLOTS], vA[SLOTS];
  int cl[SLOTS];
#pragma unroll
  for (int k = 0; k < SLOTS; k++) { vT[k] = 0.0; vA[k] = 0.0; cl[k] = S.halo + min(i, S.nrows - 1); }
  if (i < S.nrows) {
    const int g = S.row0 + i, ix = g / S.nyp, iy = g - ix * S.nyp;
    const bool wall = (ix == 0 || ix == S.nx);
    if (wall) {
      vT[4] = 1.0;   // identity row: q = 0 on x = 0, L (scft.cc:599-606)
    } else {
      const double hx = S.hx, hy = S.hy, jxw = hx * hy / 4;
      const double gm = (1.0 - 0.5773502691896257) / 2, gp = (1.0 + 0.5773502691896257) / 2;
      for (int ex = ix - 1; ex <= ix; ex++) {
        if (ex < 0 || ex >= S.nx) continue;
        for (int ey = iy - 1; ey <= iy; ey++) {
          if (ey < 0 || ey >= S.ny) continue;
          const int ax = ix - ex, ay = iy - ey;
          double en[2][2];   // eta at the element's nodes [bx][by]
#pragma unroll
          for (int bx = 0; bx < 2; bx++)
#pragma unroll
            for (int by = 0; by < 2; by++) en[bx][by] = S.eta[(size_t)(ex + bx) * S.nyp + ey + by];
#pragma unroll
          for (int bx = 0; bx < 2; bx++)
#pragma unroll
            for (int by = 0; by < 2; by++) {
              const int slot = (bx - ax + 1) * 3 + (by - ay + 1);
              const double mx = (ax == bx) ? 2.0 : 1.0, my = (ay == by) ? 2.0 : 1.0;
              const double kx = (ax == bx) ? 1.0 : -1.0, ky = (ay == by) ? 1.0 : -1.0;
              const double a = (hx / 6) * (hy / 6) * mx * my;                               // (phi_a, phi_b)
              const double b = (kx / hx) * (hy / 6) * my + (hx / 6) * mx * (ky / hy);         // (grad phi_a, grad phi_b)
              double c = 0.0;                                                                // (eta_h phi_a, phi_b), 2x2 Gauss
#pragma unroll
              for (int qx = 0; qx < 2; qx++)
#pragma unroll
                for (int qy = 0; qy < 2; qy++) {
                  const double u1 = qx ? gp : gm, u0 = 1 - u1, v1 = qy ? gp : gm, v0 = 1 - v1;   // shape values at the point
                  const double sa = (ax ? u1 : u0) * (ay ? v1 : v0), sb = (bx ? u1 : u0) * (by ? v1 : v0);
                  const double eh = en[0][0] * u0 * v0 + en[1][0] * u1 * v0 + en[0][1] * u0 * v1 + en[1][1] * u1 * v1;
                  c += sa * sb * eh * jxw;
                }
              vA[slot] += a;
              vT[slot] += a + S.dt * (b + c);
            }
        }
      }
    }
#pragma unroll
    for (int dx = -1; dx <= 1; dx++)
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
        const int slot = (dx + 1) * 3 + (dy + 1), jx = ix + dx, jy = iy + dy;
        if (jx < 0 || jx > S.nx || jy < 0 || jy >= S.nyp || jx == 0 || jx == S.nx) {
          if (!(dx == 0 && dy == 0)) { vT[slot] = 0.0; vA[slot] = 0.0; }   // outside, or a Dirichlet column
          if (dx == 0 && dy == 0) vA[slot] = wall ? 0.0 : vA[slot];
        } else {
          cl[slot] = S.halo + (jx * S.nyp + jy - S.row0);
        }
      }
    S.dinv[i] = 1.0 / vT[4];
  }
#pragma unroll
  for (int k = 0; k < SLOTS; k++) { S.col[sell(i, k)] = cl[k]; S.valT[sell(i, k)] = vT[k]; S.valA[sell(i, k)] = vA[k]; }
}

// ------------------------------------------------------------------------------------------------
// building blocks shared by the persistent kernel and the multi-GPU step kernels
struct Part { double a, b, c, d; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// block-level sum of 4 values -> partial[block][0..3]
__device__ void block_partials(Part v, double *out) {
  __shared__ double sh[4][TPB2 / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v.a = warp_sum(v.a); v.b = warp_sum(v.b); v.c = warp_sum(v.c); v.d = warp_sum(v.d);
  if (lane == 0) { sh[0][w] = v.a; sh[1][w] = v.b; sh[2][w] = v.c; sh[3][w] = v.d; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < TPB2 / 32; i++) s += sh[threadIdx.x][i];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// fixed-order sum over the blocks' partials (every block computes the same value)
__device__ Part reduce_partials(const double *partial, int nblocks) {
  Part t = {0, 0, 0, 0};
  __shared__ double tot[4];
  if (threadIdx.x < 32) {
    double a = 0, b = 0, c = 0, d = 0;
    for (int i = threadIdx.x; i < nblocks; i += 32) {
      a += partial[4 * i]; b += partial[4 * i + 1]; c += partial[4 * i + 2]; d += partial[4 * i + 3];
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
    if (threadIdx.x == 0) { tot[0] = a; tot[1] = b; tot[2] = c; tot[3] = d; }
  }
  __syncthreads();
  t.a = tot[0]; t.b = tot[1]; t.c = tot[2]; t.d = tot[3];
  __syncthreads();
  return t;
}

// Row `i` of T v (and of A v when WANT_A) WITHOUT reading the assembled matrix: the structured Q1 mesh makes the column
// pattern implicit and the element matrices closed-form, so the 108 bytes per row of stored values and column indices are
// replaced by 9 values of v and 9 values of eta, almost all of them L1/L2 hits.  Per cell (nodes [bx][by], row node a):
//   ((A + dt B)_e v)_a = sum_b e(rel(a,b)) v_b                       constant element entries, Sys2D::eA / eT
//   (C_e v)_a = (eta_h phi_a, sum_b v_b phi_b) with the 1-D triple products  int phi_a phi_b phi_k = h/4 (a=b=k), h/12 (else):
//             = hx hy/144 [ (E0+E1)(V0+V1) + 2 (eta_0a+eta_1a)(v_0a+v_1a) + 2 E_ax V_ax + 4 eta_aa v_aa ],
//     E_k = eta[k][0]+eta[k][1], V_b = v[b][0]+v[b][1]  — the exact integral, i.e. what the 2x2 Gauss rule of the assembly
//     (scft.cc:610,653-655) evaluates, so matrix-free and assembled rows agree to rounding (tests).
// Dirichlet columns: wall rows are identity rows of T and zero rows of A (scft.cc:599-606); the vectors this is applied to
// (q, z) vanish on the wall nodes, so the wall columns need no masking.
struct RowTA { double t, a; };
template <bool PEER, bool WANT_A>
__device__ __forceinline__ RowTA apply_row(const Sys2D &S, const double *__restrict__ v, int i) {
  const int g = S.row0 + i, ix = g / S.nyp, iy = g - ix * S.nyp;
  RowTA out = {0.0, 0.0};
  auto val = [&](int o) -> double {   // entry o of v (owned part starts at 0, halos on either side)
    if (PEER && (o < 0 || o >= S.nrows)) return __ldcv(v + o);
    return v[o];
  };
  if (ix == 0 || ix == S.nx) { out.t = val(i); return out; }
  const double *__restrict__ eta = S.eta + g;
#pragma unroll
  for (int ex = 0; ex < 2; ex++) {       // cell to the left (ex = 0) / right (ex = 1) of the row node
#pragma unroll
    for (int ey = 0; ey < 2; ey++) {     // cell below / above
      if ((ey == 0 && iy == 0) || (ey == 1 && iy == S.ny)) continue;
      const int dxo = (ex == 0 ? -S.nyp : S.nyp), dyo = (ey == 0 ? -1 : 1);
      // the row node is node a of the cell; xo = its x-neighbour, yo = its y-neighbour, dd = the diagonal node
      const double v_aa = val(i), v_xo = val(i + dxo), v_yo = val(i + dyo), v_dd = val(i + dxo + dyo);
      const double e_aa = eta[0], e_xo = eta[dxo], e_yo = eta[dyo], e_dd = eta[dxo + dyo];
      double t = fma(S.eT[0], v_aa, fma(S.eT[1], v_xo, fma(S.eT[2], v_yo, S.eT[3] * v_dd)));
      const double Ea = e_aa + e_yo, Eo = e_xo + e_dd, Va = v_aa + v_yo, Vo = v_xo + v_dd;   // sums over y at the row node's x / the other x
      const double c = fma(Ea + Eo, Va + Vo, fma(2.0 * (e_aa + e_xo), v_aa + v_xo, fma(2.0 * Ea, Va, 4.0 * e_aa * v_aa)));
      out.t += fma(S.cC, c, t);
      if (WANT_A) out.a += fma(S.eA[0], v_aa, fma(S.eA[1], v_xo, fma(S.eA[2], v_yo, S.eA[3] * v_dd)));
    }
  }
  return out;
}

// Initial guess of contour step j and its residual.  Step 1 starts from x0 = q_0, step 2 from the linear extrapolation
// 2 q_1 - q_0, every later step from the quadratic one x0 = 3 q_{j-1} - 3 q_{j-2} + q_{j-3} along the contour (33 -> 22 -> 18 CG
// iterations per step at 1M DOFs for none / linear / quadratic).  The residual of the guess needs no second matrix application
// and no halo of the older slices: T is linear, so T x0 = 3 T q_{j-1} - 3 T q_{j-2} + T q_{j-3}, with T q_{j-1} applied here to
// the actual q (halos at hand) and the older two kept from the previous steps (w1, w2).  (Using the CG identity T q = b - r_final instead saves the application but lets the
// drift of the recursive residual accumulate over the 2048 steps: 1.7e-9 instead of 7e-11 against the 1-D engine.)
__device__ __forceinline__ void step_guess(const Vec2D &V, int j, int i, double q, double tq, double &x0, double &tx0) {
  const double q1 = V.qp[i], t1 = V.w1[i];   // q_{j-2}, T q_{j-2}
  if (j == 1) { x0 = q; tx0 = tq; }
  else if (j == 2 || !V.qpp) { x0 = 2.0 * q - q1; tx0 = 2.0 * tq - t1; }
  else {   // quadratic: 3 q_{j-1} - 3 q_{j-2} + q_{j-3}
    x0 = 3.0 * (q - q1) + V.qpp[i]; tx0 = 3.0 * (tq - t1) + V.w2[i];
  }
  if (V.qpp) V.w2[i] = t1;
  V.w1[i] = tq;
}
// bookkeeping of the extrapolation at the end of a step, before q is overwritten by x
__device__ __forceinline__ void step_finish(const Vec2D &V, int i) {
  if (V.qpp) V.qpp[i] = V.qp[i];
  V.qp[i] = V.q[i];
}

// start of a contour step: b = A q, initial guess x0 (step_guess), r = b - T x0, z = D^-1 r, p = w = 0; partial bb = b.b
__device__ Part step_begin(const Sys2D &S, const Vec2D &V, int j, int tid, int nthreads) {
  Part acc = {0, 0, 0, 0};
  for (int i = tid; i < S.nslices * 32; i += nthreads) {
    if (i >= S.nrows) continue;
    const RowTA ta = apply_row<false, true>(S, V.q, i);   // q is stored like z (with halos), see host
    double x0, tx0;
    step_guess(V, j, i, V.q[i], ta.t, x0, tx0);
    const double bq = ta.a, r = bq - tx0;
    V.b[i] = bq; V.x[i] = x0; V.r[i] = r; V.z[i] = S.dinv[i] * r; V.p[i] = 0.0; V.w[i] = 0.0;
    acc.d = fma(bq, bq, acc.d);
  }
  return acc;
}

// entry `c` (local index incl. halo offset) of a vector with halos; PEER: halo entries are written by another GPU
// and must not be served from this SM's L1
template <bool PEER>
__device__ __forceinline__ double gather(const double *v, int c, const Sys2D &S) {
  const int o = c - S.halo;
  if (PEER && (o < 0 || o >= S.nrows)) return __ldcv(v + o);
  return v[o];
}

// s = T z; partial sums gamma = r.z, delta = z.s, rr = r.r
template <bool PEER = false>
__device__ Part cg_spmv(const Sys2D &S, const Vec2D &V, int tid, int nthreads, int phase = 0) {
  Part acc = {0, 0, 0, 0};
  // no aliasing between the gathered vector and the output: lets the loads of the next rows issue early
  const double *__restrict__ zv = V.z;
  const double *__restrict__ rv = V.r;
  double *__restrict__ sv_out = V.s;
#pragma unroll 2
  for (int i = tid; i < S.nslices * 32; i += nthreads) {
    if (i >= S.nrows) continue;
    if (phase != 0) {   // rows of the first / last owned node column are the only ones that read halo entries
      const bool edge = (i < S.halo) || (i >= S.nrows - S.halo);
      if (edge != (phase == 2)) continue;
    }
    const double sv = apply_row<PEER, false>(S, zv, i).t;
    sv_out[i] = sv;
    const double r = rv[i], z = zv[i];
    acc.a = fma(r, z, acc.a); acc.b = fma(z, sv, acc.b); acc.c = fma(r, r, acc.c);
  }
  return acc;
}

// p = z + beta p; w = s + beta w; x += alpha p; r -= alpha w; z = D^-1 r
__device__ void cg_update(const Sys2D &S, const Vec2D &V, double alpha, double beta, int tid, int nthreads) {
  double *__restrict__ pv = V.p, *__restrict__ wv = V.w, *__restrict__ xv = V.x, *__restrict__ rv = V.r, *__restrict__ zv = V.z;
  const double *__restrict__ sv = V.s, *__restrict__ dinv = S.dinv;
#pragma unroll 4
  for (int i = tid; i < S.nrows; i += nthreads) {
    const double p = fma(beta, pv[i], zv[i]), w = fma(beta, wv[i], sv[i]);
    pv[i] = p; wv[i] = w;
    xv[i] = fma(alpha, p, xv[i]);
    const double r = fma(-alpha, w, rv[i]);
    rv[i] = r; zv[i] = dinv[i] * r;
  }
}

struct March2D {
  Sys2D S; Vec2D V; Red2D R;
  int nsteps, maxit, store_full;
  double rtol;
  const double *wq;       // pair weights
  double *hist;           // [nslices_hist][nrows]
  double *phi;            // [nrows]
  long long *iters;       // total CG iterations
};

// end of a contour step: q = x, history store, fused quadrature
__device__ void step_end(const March2D &M, int j, int tid, int nthreads) {
  const int n = M.nsteps;
  const bool pairing = 2 * j > n;
  const double wj = (2 * j >= n) ? M.wq[j] : 0.0;
  for (int i = tid; i < M.S.nrows; i += nthreads) {
    const double qv = M.V.x[i];
    step_finish(M.V, i);
    M.V.q[i] = qv;
    if (M.store_full || 2 * j < n) M.hist[(size_t)j * M.S.nrows + i] = qv;
    if (2 * j >= n) {
      const double qo = pairing ? M.hist[(size_t)(n - j) * M.S.nrows + i] : qv;
      M.phi[i] = fma(wj * qv, qo, M.phi[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// single GPU: the whole march in one persistent cooperative kernel
__global__ void __launch_bounds__(TPB2, MINB2D) march2d_persistent_kernel(March2D M) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const Sys2D &S = M.S; const Vec2D &V = M.V;
  int slot = 0;
  long long iters = 0;
  for (int i = tid; i < S.nrows; i += nthreads) {   // q(x,0) = 1 inside, 0 on the walls (drivescft.cc:120-127)
    const int ix = (S.row0 + i) / S.nyp;
    const double q0 = (ix == 0 || ix == S.nx) ? 0.0 : 1.0;
    V.q[i] = q0; M.phi[i] = 0.0; M.hist[i] = q0;
  }
  grid.sync();
  for (int j = 1; j <= M.nsteps; j++) {
    Part pb = step_begin(S, V, j, tid, nthreads);
    block_partials(pb, M.R.partial + ((size_t)slot * gridDim.x + blockIdx.x) * 4);
    grid.sync();
    const double bb = reduce_partials(M.R.partial + (size_t)slot * gridDim.x * 4, gridDim.x).d;
    slot ^= 1;
    const double tol2 = M.rtol * M.rtol * bb;
    double gamma_old = 1.0, alpha_old = 1.0;
    for (int it = 0; it < M.maxit; it++) {
      Part pa = cg_spmv(S, V, tid, nthreads);
      block_partials(pa, M.R.partial + ((size_t)slot * gridDim.x + blockIdx.x) * 4);
      grid.sync();
      const Part t = reduce_partials(M.R.partial + (size_t)slot * gridDim.x * 4, gridDim.x);
      slot ^= 1;
      if (t.c <= tol2) break;                                  // ||r|| <= rtol ||b||
      const double beta = (it == 0) ? 0.0 : t.a / gamma_old;
      const double alpha = (it == 0) ? t.a / t.b : t.a / (t.b - beta * t.a / alpha_old);
      cg_update(S, V, alpha, beta, tid, nthreads);
      gamma_old = t.a; alpha_old = alpha;
      iters++;
      grid.sync();
    }
    step_end(M, j, tid, nthreads);
    grid.sync();
  }
  if (tid == 0) *M.iters = iters;
}

// ------------------------------------------------------------------------------------------------
// multi GPU, peer-memory version: ONE persistent cooperative kernel per rank for the whole march.  Halo
// columns are stored straight into the neighbours' vectors over NVLink by the threads that produce them,
// and the dot products are all-reduced through peer-mapped slots: compute and exchange live in the same
// kernel, nothing is launched per iteration.  (cudaIpc handles of the exchange buffers travel through
// torch.distributed; see scftb2d_p2p_handle / scftb2d_p2p_attach.)
typedef unsigned long long ull;
struct P2P {
  int rank, world;
  double *const *peers;      // [world] base of every rank's exchange buffer (own entry = local pointer)
  size_t off_q, off_z;       // offsets (doubles) of the q / z vectors (halo-left start) inside an exchange buffer
  size_t off_red;            // [2][world][4] reduction slots
  size_t off_flag;           // ull flags: [0] left-in, [1] right-in, [2 + par*world + r] reduction arrivals
  int nrows_left;            // owned rows of the left neighbour
  long long timeout;         // cycles
  volatile long long *dbg;   // host-mapped: [0] stage of a timed-out wait, [1] expected seq, [2] seen value, [3] block
};

__device__ __forceinline__ ull ld_acquire_sys(const ull *p) {
  ull v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(ull *p, ull v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ ull ld_acquire_gpu(const ull *p) {
  ull v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(ull *p, ull v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// spin (acquire, system scope) until a flag written by a peer GPU reaches seq
__device__ __forceinline__ void wait_flag(const ull *f, ull seq, const P2P &X, int stage) {
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < seq)
    if (clock64() - t0 > X.timeout) {   // a peer never arrived: abort this kernel instead of hanging the GPU
      X.dbg[0] = stage; X.dbg[1] = (long long)seq; X.dbg[2] = (long long)ld_acquire_sys(f); X.dbg[3] = blockIdx.x;
      __threadfence_system();
      __trap();
    }
}

// boundary entries of an owned vector go straight into the neighbours' halo regions
__device__ __forceinline__ bool push_boundary(const Sys2D &S, const P2P &X, size_t off_vec, int i, double v) {
  bool pushed = false;
  if (X.rank > 0 && i < S.halo) { X.peers[X.rank - 1][off_vec + S.halo + X.nrows_left + i] = v; pushed = true; }
  if (X.rank + 1 < X.world && i >= S.nrows - S.halo) { X.peers[X.rank + 1][off_vec + (i - (S.nrows - S.halo))] = v; pushed = true; }
  return pushed;
}

// after a grid barrier that follows the pushes: raise the neighbours' flags ...
__device__ __forceinline__ void halo_signal(const P2P &X, ull seq) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // the pushing threads fenced (system scope) before the grid barrier
    if (X.rank > 0) st_release_sys((ull *)(X.peers[X.rank - 1] + X.off_flag) + 1, seq);
    if (X.rank + 1 < X.world) st_release_sys((ull *)(X.peers[X.rank + 1] + X.off_flag) + 0, seq);
  }
}
// ... and, as late as possible (only the rows next to a slab boundary read halo entries), wait for ours
__device__ void halo_wait(const P2P &X, ull seq) {
  ull *myflags = (ull *)(X.peers[X.rank] + X.off_flag);
  if (threadIdx.x == 0) {
    if (X.rank > 0) wait_flag(myflags + 0, seq, X, 1);
    if (X.rank + 1 < X.world) wait_flag(myflags + 1, seq, X, 2);
  }
  __syncthreads();   // halo entries are read with ld.cv (gather<true>), so no L1 invalidation is needed here
}

// all-reduce of the block partials over the ranks; every block of every rank returns the same sums.
// Block 0 folds the partials, stores them into every peer's slot (release, system scope), gathers the peers'
// contributions in rank order and publishes the totals to the other blocks through a local flag (gpu scope):
// only one block per GPU polls peer-written memory.
__device__ Part allreduce_partials(const P2P &X, const double *partial, int nblocks, ull rseq, cg::grid_group &grid) {
  const int par = (int)(rseq & 1);
  double *self = X.peers[X.rank];
  ull *flags = (ull *)(self + X.off_flag);
  double *gsum = self + X.off_red + (size_t)2 * X.world * 4 + par * 4;   // local totals [2][4]
  ull *gready = flags + 2 + 2 * X.world;                                  // local "totals ready" sequence number
  __shared__ double tot[4];
  if (blockIdx.x == 0) {
    Part t = reduce_partials(partial, nblocks);
    if (threadIdx.x < X.world) {
      double *slot = X.peers[threadIdx.x] + X.off_red + (size_t)(par * X.world + X.rank) * 4;
      slot[0] = t.a; slot[1] = t.b; slot[2] = t.c; slot[3] = t.d;
      st_release_sys((ull *)(X.peers[threadIdx.x] + X.off_flag) + 2 + par * X.world + X.rank, rseq);
    }
    if (threadIdx.x == 0) {
      const double *sl = self + X.off_red + (size_t)par * X.world * 4;
      double a = 0, b = 0, c = 0, d = 0;
      for (int r = 0; r < X.world; r++) {   // fixed rank order: identical result everywhere
        wait_flag(flags + 2 + par * X.world + r, rseq, X, 10 + r);
        a += __ldcv(sl + 4 * r); b += __ldcv(sl + 4 * r + 1); c += __ldcv(sl + 4 * r + 2); d += __ldcv(sl + 4 * r + 3);
      }
      gsum[0] = a; gsum[1] = b; gsum[2] = c; gsum[3] = d;
      st_release_gpu(gready, rseq);
      tot[0] = a; tot[1] = b; tot[2] = c; tot[3] = d;
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(gready) < rseq)
      if (clock64() - t0 > 2 * X.timeout) { X.dbg[0] = 99; X.dbg[1] = (long long)rseq; X.dbg[3] = blockIdx.x; __threadfence_system(); __trap(); }
    tot[0] = __ldcg(gsum); tot[1] = __ldcg(gsum + 1); tot[2] = __ldcg(gsum + 2); tot[3] = __ldcg(gsum + 3);
  }
  __syncthreads();
  Part t = {tot[0], tot[1], tot[2], tot[3]};
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(TPB2, 4) march2d_p2p_kernel(March2D M, P2P X, ull seq0, ull rseq0) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const Sys2D &S = M.S; const Vec2D &V = M.V;
  ull seq = seq0, rseq = rseq0;   // halo / reduction sequence numbers continue across launches
  int slot = 0;
  long long iters = 0;
  for (int i = tid; i < S.nrows; i += nthreads) {
    const int ix = (S.row0 + i) / S.nyp;
    const double q0 = (ix == 0 || ix == S.nx) ? 0.0 : 1.0;
    V.q[i] = q0; M.phi[i] = 0.0; M.hist[i] = q0;
    if (push_boundary(S, X, X.off_q, i, q0)) __threadfence_system();   // peer stores ordered before the flag
  }
  grid.sync();
  halo_signal(X, ++seq);
  halo_wait(X, seq);
  for (int j = 1; j <= M.nsteps; j++) {
    // b = A q, x = q, r = b - T x, z = D^-1 r (boundary z goes to the neighbours at once)
    Part pb = {0, 0, 0, 0};
    for (int i = tid; i < S.nslices * 32; i += nthreads) {
      if (i >= S.nrows) continue;
      const RowTA ta = apply_row<true, true>(S, V.q, i);
      double x0, tx0;
      step_guess(V, j, i, V.q[i], ta.t, x0, tx0);
      const double bq = ta.a, r = bq - tx0, z = S.dinv[i] * r;
      V.b[i] = bq; V.x[i] = x0; V.r[i] = r; V.z[i] = z; V.p[i] = 0.0; V.w[i] = 0.0;
      if (push_boundary(S, X, X.off_z, i, z)) __threadfence_system();
      pb.d = fma(bq, bq, pb.d);
    }
    block_partials(pb, M.R.partial + ((size_t)slot * gridDim.x + blockIdx.x) * 4);
    grid.sync();
    const double bb = allreduce_partials(X, M.R.partial + (size_t)slot * gridDim.x * 4, gridDim.x, ++rseq, grid).d;
    slot ^= 1;
    halo_signal(X, ++seq);   // z of step_begin was pushed before the barrier above
    const double tol2 = M.rtol * M.rtol * bb;
    double gamma_old = 1.0, alpha_old = 1.0;
    for (int it = 0; it < M.maxit; it++) {
      // rows that touch no halo entry first; the neighbours' columns are awaited only for the two boundary columns
      Part pa = cg_spmv<true>(S, V, tid, nthreads, 1);
      halo_wait(X, seq);
      Part pb2 = cg_spmv<true>(S, V, tid, nthreads, 2);
      pa.a += pb2.a; pa.b += pb2.b; pa.c += pb2.c;
      block_partials(pa, M.R.partial + ((size_t)slot * gridDim.x + blockIdx.x) * 4);
      grid.sync();
      const Part t = allreduce_partials(X, M.R.partial + (size_t)slot * gridDim.x * 4, gridDim.x, ++rseq, grid);
      slot ^= 1;
      if (t.c <= tol2) break;
      const double beta = (it == 0) ? 0.0 : t.a / gamma_old;
      const double alpha = (it == 0) ? t.a / t.b : t.a / (t.b - beta * t.a / alpha_old);
      for (int i = tid; i < S.nrows; i += nthreads) {   // cg_update + push of the new boundary z
        const double p = fma(beta, V.p[i], V.z[i]), w = fma(beta, V.w[i], V.s[i]);
        V.p[i] = p; V.w[i] = w;
        V.x[i] = fma(alpha, p, V.x[i]);
        const double r = fma(-alpha, w, V.r[i]);
        const double z = S.dinv[i] * r;
        V.r[i] = r; V.z[i] = z;
        if (push_boundary(S, X, X.off_z, i, z)) __threadfence_system();
      }
      gamma_old = t.a; alpha_old = alpha;
      iters++;
      grid.sync();
      halo_signal(X, ++seq);
    }
    // step end: q = x (+ push), history, fused quadrature
    {
      const int n = M.nsteps;
      const bool pairing = 2 * j > n;
      const double wj = (2 * j >= n) ? M.wq[j] : 0.0;
      for (int i = tid; i < S.nrows; i += nthreads) {
        const double qv = V.x[i];
        step_finish(V, i);
        V.q[i] = qv;
        if (push_boundary(S, X, X.off_q, i, qv)) __threadfence_system();
        if (M.store_full || 2 * j < n) M.hist[(size_t)j * S.nrows + i] = qv;
        if (2 * j >= n) {
          const double qo = pairing ? M.hist[(size_t)(n - j) * S.nrows + i] : qv;
          M.phi[i] = fma(wj * qv, qo, M.phi[i]);
        }
      }
    }
    grid.sync();
    halo_signal(X, ++seq);
    halo_wait(X, seq);
  }
  if (tid == 0) { *M.iters = iters; M.R.scal[30] = (double)seq; M.R.scal[31] = (double)rseq; }
}

// ------------------------------------------------------------------------------------------------
// multi GPU: the same pieces as short kernels; scalars live in R.scal (all-reduced by NCCL in between)
__global__ void init2d_kernel(March2D M) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int i = tid; i < M.S.nrows; i += nthreads) {
    const int ix = (M.S.row0 + i) / M.S.nyp;
    const double q0 = (ix == 0 || ix == M.S.nx) ? 0.0 : 1.0;
    M.V.q[i] = q0; M.phi[i] = 0.0; M.hist[i] = q0;
  }
  if (tid == 0) *M.iters = 0;
}
__global__ void __launch_bounds__(TPB2) step_begin_kernel(March2D M, int j) {   // needs q halos; leaves bb partials
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  Part pb = step_begin(M.S, M.V, j, tid, nthreads);
  block_partials(pb, M.R.partial + (size_t)blockIdx.x * 4);
}
__global__ void __launch_bounds__(TPB2) spmv_kernel(March2D M, int par) {   // needs z halos; leaves gamma, delta, rr partials
  if (M.R.scal[8 + 8 * par + 2] != 0.0) return;                              // converged: nothing to do
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  Part pa = cg_spmv(M.S, M.V, tid, nthreads);
  block_partials(pa, M.R.partial + (size_t)blockIdx.x * 4);
}
// one block: partials -> scal[0..3] (local sums; NCCL all-reduces them in place afterwards)
__global__ void fold_partials_kernel(March2D M, int nblocks, int par) {
  if (par >= 0 && M.R.scal[8 + 8 * par + 2] != 0.0) { if (threadIdx.x < 4) M.R.scal[threadIdx.x] = 0.0; return; }
  Part t = reduce_partials(M.R.partial, nblocks);
  if (threadIdx.x == 0) { M.R.scal[0] = t.a; M.R.scal[1] = t.b; M.R.scal[2] = t.c; M.R.scal[3] = t.d; }
}
__global__ void begin_finish_kernel(March2D M) {   // after the all-reduce of bb: reset state record 0
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double *st = M.R.scal + 8;
    st[0] = 1.0; st[1] = 1.0; st[2] = 0.0; st[3] = M.R.scal[3]; st[4] = 0.0;
  }
}
// after the all-reduce of gamma, delta, rr: reads state record `par`, writes record `par^1` (no grid barrier needed)
__global__ void __launch_bounds__(TPB2) update_kernel(March2D M, int par) {
  const double *st = M.R.scal + 8 + 8 * par;
  double *sn = M.R.scal + 8 + 8 * (par ^ 1);
  const double gam = M.R.scal[0], del = M.R.scal[1], rr = M.R.scal[2], bb = st[3];
  const bool first = (st[4] == 0.0);
  const bool fin = (st[2] != 0.0) || (rr <= M.rtol * M.rtol * bb);
  const double beta = first ? 0.0 : gam / st[0];
  const double alpha = first ? gam / del : gam / (del - beta * gam / st[1]);
  if (!fin) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    cg_update(M.S, M.V, alpha, beta, tid, nthreads);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sn[3] = bb;
    if (fin) { sn[0] = st[0]; sn[1] = st[1]; sn[2] = 1.0; sn[4] = st[4]; }
    else { sn[0] = gam; sn[1] = alpha; sn[2] = 0.0; sn[4] = st[4] + 1.0; *M.iters += 1; }
  }
}
__global__ void __launch_bounds__(TPB2) step_end_kernel(March2D M, int j) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  step_end(M, j, tid, nthreads);
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time dependency on it (single-GPU use needs none)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  bool load() {
    if (h) return true;
    for (const char *nm : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
#define SYM(f, n) f = (decltype(f))dlsym(h, n); if (!f) return false
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
static Nccl g_nccl;
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;   // ncclFloat64, ncclSum (nccl.h)

}  // namespace scftb

// exchange buffers exported over cudaIpc are recycled, never freed: a peer process may still hold a mapping, and
// re-exporting memory freed and re-allocated inside the same CUDA memory block is not reliable
struct XchgArena { int device; double *base; size_t doubles; bool busy; };
static std::vector<XchgArena> g_arenas;

struct scftb2d_engine {
  scftb2d_config cfg;
  int nyp, ndof, ix0, ix1, nrows, halo, nslices, grid_persist, grid_step;
  cudaStream_t stream;
  ncclComm_t comm;
  March2D M;
  double *d_eta, *d_qbuf, *d_zbuf, *d_out, *d_f0;
  std::vector<double> h_f0x, h_w;
  long long last_iters;
  double last_ms;
  cudaGraphExec_t graph_exec;
  double *h_flag;   // pinned
  // peer-memory exchange (world > 1)
  double *d_xchg;
  size_t xchg_doubles;
  P2P p2p;
  bool attached;
  std::vector<void *> opened;
  double **d_peers;
  unsigned long long seq, rseq;
  volatile long long *h_dbg = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define NK(call)                                                                                              \
  do {                                                                                                        \
    int _r = (call);                                                                                          \
    if (_r != 0) return fail(SCFTB_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(_r));           \
  } while (0)

extern "C" {

int scftb2d_nccl_unique_id(char *id128) {
  if (!g_nccl.load()) return fail(SCFTB_ERR_STATE, "libnccl.so.2 not found");
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return SCFTB_OK;
}

int scftb2d_destroy(scftb2d_engine *e) {
  if (!e) return SCFTB_OK;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  for (void *p : e->opened) cudaIpcCloseMemHandle(p);
  for (auto &a : g_arenas) if (a.base == e->d_xchg) a.busy = false;
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  if (e->h_flag) cudaFreeHost(e->h_flag);
  if (e->h_dbg) cudaFreeHost((void *)e->h_dbg);
  if (e->ev0) { cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1); }
  if (e->comm) g_nccl.CommDestroy(e->comm);
  for (void *p : {(void *)e->M.S.col, (void *)e->M.S.valT, (void *)e->M.S.valA, (void *)e->M.S.dinv, (void *)e->d_eta,
                  (void *)e->d_peers, (void *)e->M.V.x, (void *)e->M.V.r, (void *)e->M.V.s, (void *)e->M.V.p,
                  (void *)e->M.V.w, (void *)e->M.V.b, (void *)e->M.V.qp, (void *)e->M.V.w1, (void *)e->M.V.qpp, (void *)e->M.V.w2, (void *)e->M.R.partial, (void *)e->M.R.scal, (void *)e->M.hist,
                  (void *)e->M.phi, (void *)e->M.iters, (void *)e->d_out, (void *)e->d_f0, (void *)e->M.wq})
    if (p) cudaFree(p);
  cudaStreamDestroy(e->stream);
  delete e;
  return SCFTB_OK;
}

int scftb2d_create(const scftb2d_config *cfg, const char *nccl_id128, scftb2d_engine **out) {
  if (!cfg || !out) return fail(SCFTB_ERR_ARG, "null argument");
  if (cfg->nx < 2 || cfg->ny < 1 || cfg->nsteps < 2 || cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
    return fail(SCFTB_ERR_ARG, "bad 2-D configuration");
  if (cfg->world > 1 && (cfg->nx + 1) / cfg->world < 2) return fail(SCFTB_ERR_ARG, "too few node columns per rank");
  scftb2d_engine *e = new scftb2d_engine();
  memset(&e->M, 0, sizeof(e->M));
  e->cfg = *cfg; e->comm = nullptr; e->d_eta = e->d_qbuf = e->d_zbuf = e->d_out = e->d_f0 = nullptr; e->last_iters = 0; e->last_ms = 0; e->graph_exec = nullptr; e->h_flag = nullptr; e->d_xchg = nullptr; e->attached = false; e->d_peers = nullptr; e->seq = 0; e->rseq = 0;
  const int nx = cfg->nx, ny = cfg->ny, nyp = ny + 1, n = cfg->nsteps;
  e->nyp = nyp; e->ndof = (nx + 1) * nyp;
  // slab partition: node columns [ix0, ix1) (1-D slab along x, one node column of halo per neighbour)
  e->ix0 = (int)((long long)(nx + 1) * cfg->rank / cfg->world);
  e->ix1 = (int)((long long)(nx + 1) * (cfg->rank + 1) / cfg->world);
  e->nrows = (e->ix1 - e->ix0) * nyp;
  e->halo = cfg->world > 1 ? nyp : 0;
  e->nslices = (e->nrows + 31) / 32;
  if (cfg->quadrature == SCFTB_QUAD_ROMBERG) {
    if (romberg_weights(n, 1.0 / n, e->h_w)) { delete e; return fail(SCFTB_ERR_ARG, "Romberg needs nsteps = 2^k >= 16"); }
  } else trapezoid_weights(n, 1.0 / n, e->h_w);
  std::vector<double> wq(n + 1);
  for (int j = 0; j <= n; j++) wq[j] = (2 * j > n) ? 2.0 * e->h_w[j] : ((2 * j == n) ? e->h_w[j] : 0.0);
  std::vector<double> xs(nx + 1);
  for (int i = 0; i <= nx; i++) xs[i] = cfg->L * i / nx;
  e->h_f0x.resize(nx + 1);
  f0_given(nx + 1, xs.data(), cfg->tau, e->h_f0x.data());
#define CK2(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { int rc = fail(SCFTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); scftb2d_destroy(e); return rc; } } while (0)
  CK2(cudaSetDevice(cfg->device));
  CK2(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  const size_t nr = e->nrows, nv = (size_t)e->nslices * SLOTS * 32, nh = nr + 2 * (size_t)e->halo;
  Sys2D &S = e->M.S;
  S.nx = nx; S.ny = ny; S.nyp = nyp; S.row0 = e->ix0 * nyp; S.nrows = e->nrows; S.halo = e->halo; S.nslices = e->nslices;
  S.hx = cfg->L / nx; S.hy = cfg->Ly / ny; S.dt = 1.0 / n;
  {
    const double hx = S.hx, hy = S.hy, m = hx * hy / 36.0;
    const double eB[4] = {hy / (3 * hx) + hx / (3 * hy), -hy / (3 * hx) + hx / (6 * hy), hy / (6 * hx) - hx / (3 * hy), -hy / (6 * hx) - hx / (6 * hy)};
    const double mA[4] = {4 * m, 2 * m, 2 * m, m};
    for (int k = 0; k < 4; k++) { S.eA[k] = mA[k]; S.eT[k] = mA[k] + S.dt * eB[k]; }
    S.cC = S.dt * hx * hy / 144.0;
  }
  CK2(cudaMalloc(&S.col, sizeof(int) * nv));
  CK2(cudaMalloc(&S.valT, sizeof(double) * nv));
  CK2(cudaMalloc(&S.valA, sizeof(double) * nv));
  CK2(cudaMalloc(&S.dinv, sizeof(double) * nr));
  CK2(cudaMalloc(&e->d_eta, sizeof(double) * e->ndof));
  S.eta = e->d_eta;
  {  // one exportable allocation: q and z (with halos), the reduction slots and the flags
    // the layout must be identical on every rank (peers address each other's buffers with their own offsets):
    // size the vectors for the widest slab
    const size_t nh_max = (size_t)((nx + 1 + cfg->world - 1) / cfg->world + 1) * nyp + 2 * (size_t)e->halo;
    const size_t nhp = (nh_max + 15) / 16 * 16, red = (size_t)2 * cfg->world * 4 + 8, fl = (size_t)(3 + 2 * cfg->world + 14) / 16 * 16 + 16;
    e->p2p.off_q = 0; e->p2p.off_z = nhp; e->p2p.off_red = 2 * nhp; e->p2p.off_flag = 2 * nhp + (red + 15) / 16 * 16;
    e->xchg_doubles = e->p2p.off_flag + fl;
    for (auto &a : g_arenas)
      if (!a.busy && a.device == cfg->device && a.doubles >= e->xchg_doubles) { e->d_xchg = a.base; a.busy = true; break; }
    if (!e->d_xchg) {
      const size_t want = std::max(e->xchg_doubles, (size_t)1 << 20);
      CK2(cudaMalloc(&e->d_xchg, sizeof(double) * want));
      g_arenas.push_back({cfg->device, e->d_xchg, want, true});
    }
    CK2(cudaMemset(e->d_xchg, 0, sizeof(double) * e->xchg_doubles));
    e->d_qbuf = e->d_xchg + e->p2p.off_q; e->d_zbuf = e->d_xchg + e->p2p.off_z;
    e->p2p.rank = cfg->rank; e->p2p.world = cfg->world; e->p2p.peers = nullptr;
    const int lx0 = cfg->rank > 0 ? (int)((long long)(nx + 1) * (cfg->rank - 1) / cfg->world) : 0;
    e->p2p.nrows_left = cfg->rank > 0 ? (e->ix0 - lx0) * nyp : 0;
    // bound on a peer-flag wait, in SM clocks (~2 GHz).  Ranks can be skewed by seconds at the first launch (module load,
    // host paging, a large eta upload on one rank), so the default is a minute, not a few launch latencies;
    // SCFTB_P2P_TIMEOUT_S overrides it.  Expiry still __trap()s (a peer that never arrives would otherwise hang the GPU).
    {
      const char *ts = getenv("SCFTB_P2P_TIMEOUT_S");
      double sec = ts ? atof(ts) : 60.0;
      if (!(sec > 0.0)) sec = 60.0;
      e->p2p.timeout = (long long)(sec * 2.0e9);
    }
    CK2(cudaHostAlloc((void **)&e->h_dbg, sizeof(long long) * 8, cudaHostAllocMapped));
    memset((void *)e->h_dbg, 0, sizeof(long long) * 8);
    long long *ddbg = nullptr;
    CK2(cudaHostGetDevicePointer((void **)&ddbg, (void *)e->h_dbg, 0));
    e->p2p.dbg = ddbg;
  }
  Vec2D &V = e->M.V;
  V.q = e->d_qbuf + e->halo; V.z = e->d_zbuf + e->halo;
  for (double **p : {&V.x, &V.r, &V.s, &V.p, &V.w, &V.b, &V.qp, &V.w1}) CK2(cudaMalloc(p, sizeof(double) * nr));
  V.qpp = V.w2 = nullptr;
  // quadratic extrapolation of the CG start by default; SCFTB_2D_LINEAR=1: linear only (A/B runs)
  if (!getenv("SCFTB_2D_LINEAR")) {
    CK2(cudaMalloc(&V.qpp, sizeof(double) * nr));
    CK2(cudaMalloc(&V.w2, sizeof(double) * nr));
  }
  e->M.nsteps = n; e->M.maxit = cfg->maxit > 0 ? cfg->maxit : 100000; e->M.rtol = cfg->rtol > 0 ? cfg->rtol : 1e-12;
  e->M.store_full = cfg->store_history;
  const size_t nsl = cfg->store_history ? n + 1 : n / 2 + 1;
  CK2(cudaMalloc(&e->M.hist, sizeof(double) * nsl * nr));
  CK2(cudaMalloc(&e->M.phi, sizeof(double) * nr));
  CK2(cudaMalloc(&e->M.iters, sizeof(long long)));
  CK2(cudaMalloc(&e->d_out, sizeof(double) * nr));
  CK2(cudaMalloc(&e->d_f0, sizeof(double) * (nx + 1)));
  CK2(cudaMemcpy(e->d_f0, e->h_f0x.data(), sizeof(double) * (nx + 1), cudaMemcpyHostToDevice));
  double *dwq = nullptr;
  CK2(cudaMalloc(&dwq, sizeof(double) * (n + 1)));
  CK2(cudaMemcpy(dwq, wq.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
  e->M.wq = dwq;
  int sms = 0, occ = 0;
  CK2(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
  CK2(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, march2d_persistent_kernel, TPB2, 0));
  {
    int occp = 0;
    CK2(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, march2d_p2p_kernel, TPB2, 0));
    if (cfg->world > 1) occ = std::min(occ, occp);
  }
  {
    const char *cap = getenv("SCFTB_2D_BLOCKS_PER_SM");   // tuning knob; default 8 (memory-latency bound: more warps in flight)
    occ = std::max(1, std::min(occ, cap ? atoi(cap) : 8));
  }
  e->grid_persist = std::max(1, std::min(sms * occ, (e->nrows + 4 * TPB2 - 1) / (4 * TPB2)));   // small meshes: fewer blocks, cheaper barriers
  int occ2 = 0;
  CK2(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, spmv_kernel, TPB2, 0));
  e->grid_step = sms * std::max(1, std::min(occ2, 4));
  const int gmax = std::max(e->grid_persist, e->grid_step);
  CK2(cudaMalloc(&e->M.R.partial, sizeof(double) * 2 * gmax * 4));
  CK2(cudaMalloc(&e->M.R.scal, sizeof(double) * 32));
  CK2(cudaMemset(e->M.R.scal, 0, sizeof(double) * 32));
  CK2(cudaMallocHost(&e->h_flag, sizeof(double) * 4));
  if (cfg->world > 1) {
    if (!nccl_id128) { scftb2d_destroy(e); return fail(SCFTB_ERR_ARG, "world > 1 needs the NCCL unique id of rank 0"); }
    if (!g_nccl.load()) { scftb2d_destroy(e); return fail(SCFTB_ERR_STATE, "libnccl.so.2 not found"); }
    ncclUniqueId id;
    memcpy(id.internal, nccl_id128, 128);
    int r = g_nccl.CommInitRank(&e->comm, cfg->world, id, cfg->rank);
    if (r != 0) { int rc = fail(SCFTB_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); scftb2d_destroy(e); return rc; }
  }
  CK2(cudaDeviceSynchronize());   // the zeroed exchange buffer is in place before any peer can be told about it
  *out = e;
  return SCFTB_OK;
}

// cudaIpc handle (64 bytes) of this rank's exchange buffer
int scftb2d_p2p_handle(scftb2d_engine *e, char *handle64) {
  if (!e || !handle64) return fail(SCFTB_ERR_ARG, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->d_xchg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return SCFTB_OK;
}

// handles: world x 64 bytes, entry r = scftb2d_p2p_handle of rank r.  From then on scftb2d_residual runs the
// peer-memory persistent kernel instead of the NCCL path.
int scftb2d_p2p_attach(scftb2d_engine *e, const char *handles) {
  if (!e || !handles) return fail(SCFTB_ERR_ARG, "null argument");
  if (e->cfg.world < 2) return fail(SCFTB_ERR_STATE, "p2p needs world > 1");
  CK(cudaSetDevice(e->cfg.device));
  std::vector<double *> peers(e->cfg.world, nullptr);
  for (int r = 0; r < e->cfg.world; r++) {
    if (r == e->cfg.rank) { peers[r] = e->d_xchg; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)64 * r, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->opened.push_back(p);
    peers[r] = (double *)p;
  }
  CK(cudaMalloc(&e->d_peers, sizeof(double *) * e->cfg.world));
  CK(cudaMemcpy(e->d_peers, peers.data(), sizeof(double *) * e->cfg.world, cudaMemcpyHostToDevice));
  e->p2p.peers = e->d_peers;
  e->attached = true;
  return SCFTB_OK;
}

// close the peer mappings (call on every rank, then synchronise the ranks, then destroy: exported memory must
// not be freed while a peer still maps it)
int scftb2d_p2p_detach(scftb2d_engine *e) {
  if (!e) return fail(SCFTB_ERR_ARG, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->stream));
  for (void *p : e->opened) cudaIpcCloseMemHandle(p);
  e->opened.clear();
  e->attached = false;
  return SCFTB_OK;
}

// exchange the one-column halos of v (device pointer to the owned part; halos live just outside it)
static int halo_exchange(scftb2d_engine *e, double *v) {
  const int r = e->cfg.rank, w = e->cfg.world, h = e->halo, nr = e->nrows;
  NK(g_nccl.GroupStart());
  if (r > 0) { NK(g_nccl.Send(v, h, NCCL_DOUBLE, r - 1, e->comm, e->stream)); NK(g_nccl.Recv(v - h, h, NCCL_DOUBLE, r - 1, e->comm, e->stream)); }
  if (r + 1 < w) { NK(g_nccl.Send(v + nr - h, h, NCCL_DOUBLE, r + 1, e->comm, e->stream)); NK(g_nccl.Recv(v + nr, h, NCCL_DOUBLE, r + 1, e->comm, e->stream)); }
  NK(g_nccl.GroupEnd());
  return SCFTB_OK;
}

// eta: field on ALL (nx+1)(ny+1) nodes (host), out: sign*(phi0 - phi) on this rank's rows (host, nrows values)
int scftb2d_residual(scftb2d_engine *e, const double *eta, double *out) {
  if (!e || !eta || !out) return fail(SCFTB_ERR_ARG, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = e->stream;
  CK(cudaMemcpyAsync(e->d_eta, eta, sizeof(double) * e->ndof, cudaMemcpyHostToDevice, st));
  assemble2d_kernel<<<(e->nslices * 32 + 255) / 256, 256, 0, st>>>(e->M.S);
  g_launches++;
  if (!e->ev0) { CK(cudaEventCreate(&e->ev0)); CK(cudaEventCreate(&e->ev1)); }   // owned by the engine: no leak on error returns
  cudaEvent_t e0 = e->ev0, e1 = e->ev1;
  CK(cudaEventRecord(e0, st));
  March2D M = e->M;
  if (e->cfg.world > 1 && e->attached) {
    P2P X = e->p2p;
    unsigned long long s0 = e->seq, r0 = e->rseq;
    void *args[] = {&M, &X, &s0, &r0};
    CK(cudaLaunchCooperativeKernel((void *)march2d_p2p_kernel, dim3(e->grid_persist), dim3(TPB2), args, 0, st));
    g_launches++;
    double sq[2];
    cudaError_t ce = cudaMemcpyAsync(sq, M.R.scal + 30, sizeof(double) * 2, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) {
      char msg[256];
      snprintf(msg, sizeof msg, "peer-memory march failed (%s); timed-out wait: stage %lld expected seq %lld saw %lld in block %lld (rank %d)",
               cudaGetErrorString(ce), e->h_dbg[0], e->h_dbg[1], e->h_dbg[2], e->h_dbg[3], e->cfg.rank);
      return fail(SCFTB_ERR_CUDA, msg);
    }
    e->seq = (unsigned long long)sq[0]; e->rseq = (unsigned long long)sq[1];
  } else if (e->cfg.world == 1) {
    void *args[] = {&M};
    CK(cudaLaunchCooperativeKernel((void *)march2d_persistent_kernel, dim3(e->grid_persist), dim3(TPB2), args, 0, st));
    g_launches++;
  } else {
    const int G = e->grid_step;
    init2d_kernel<<<G, TPB2, 0, st>>>(M);
    g_launches++;
    // 8 CG iterations (halo exchange, SpMV + partial dots, fold, all-reduce, update) captured once as a CUDA
    // graph and replayed; the device-side convergence flag is polled after every replay
    constexpr int ITERS_PER_GRAPH = 8;
    if (!e->graph_exec) {
      cudaGraph_t graph;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      for (int it = 0; it < ITERS_PER_GRAPH; it++) {
        const int par = it & 1;
        int rc = halo_exchange(e, M.V.z);
        if (rc) { cudaStreamEndCapture(st, &graph); return rc; }
        spmv_kernel<<<G, TPB2, 0, st>>>(M, par);
        fold_partials_kernel<<<1, 32, 0, st>>>(M, G, par);
        NK(g_nccl.AllReduce(M.R.scal, M.R.scal, 3, NCCL_DOUBLE, NCCL_SUM, e->comm, st));
        update_kernel<<<G, TPB2, 0, st>>>(M, par);
      }
      CK(cudaStreamEndCapture(st, &graph));
      CK(cudaGraphInstantiate(&e->graph_exec, graph, 0));
      CK(cudaGraphDestroy(graph));
    }
    double *flag = e->h_flag;
    for (int j = 1; j <= M.nsteps; j++) {
      int rc = halo_exchange(e, M.V.q);
      if (rc) return rc;
      step_begin_kernel<<<G, TPB2, 0, st>>>(M, j);
      fold_partials_kernel<<<1, 32, 0, st>>>(M, G, -1);
      NK(g_nccl.AllReduce(M.R.scal, M.R.scal, 4, NCCL_DOUBLE, NCCL_SUM, e->comm, st));
      begin_finish_kernel<<<1, 32, 0, st>>>(M);
      g_launches += 3;
      bool conv = false;
      for (int it = 0; it < M.maxit && !conv; it += ITERS_PER_GRAPH) {
        CK(cudaGraphLaunch(e->graph_exec, st));
        g_launches += 3 * ITERS_PER_GRAPH;
        CK(cudaMemcpyAsync(flag, M.R.scal + 8 + 2, sizeof(double), cudaMemcpyDeviceToHost, st));   // record 0 (even count)
        CK(cudaStreamSynchronize(st));
        conv = flag[0] != 0.0;
      }
      step_end_kernel<<<G, TPB2, 0, st>>>(M, j);
      g_launches++;
    }
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  e->last_ms = ms;
  CK(cudaMemcpy(&e->last_iters, M.iters, sizeof(long long), cudaMemcpyDeviceToHost));
  std::vector<double> phi(e->nrows);
  CK(cudaMemcpy(phi.data(), M.phi, sizeof(double) * e->nrows, cudaMemcpyDeviceToHost));
  for (int i = 0; i < e->nrows; i++) {
    const int ix = e->ix0 + i / e->nyp;
    out[i] = (ix == 0 || ix == e->cfg.nx) ? 0.0 : e->cfg.sign * (e->h_f0x[ix] - phi[i]);
  }
  return SCFTB_OK;
}

int scftb2d_rows(scftb2d_engine *e, int *row0, int *nrows) {
  if (!e) return fail(SCFTB_ERR_ARG, "null engine");
  if (row0) *row0 = e->ix0 * e->nyp;
  if (nrows) *nrows = e->nrows;
  return SCFTB_OK;
}

int scftb2d_get_phi(scftb2d_engine *e, double *phi) {
  if (!e || !phi) return fail(SCFTB_ERR_ARG, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaMemcpy(phi, e->M.phi, sizeof(double) * e->nrows, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

int scftb2d_get_stats(scftb2d_engine *e, long long *cg_iterations, double *march_ms) {
  if (!e) return fail(SCFTB_ERR_ARG, "null engine");
  if (cg_iterations) *cg_iterations = e->last_iters;
  if (march_ms) *march_ms = e->last_ms;
  return SCFTB_OK;
}

// CSR view of this rank's rows of T and A (global column indices), for inspection and tests.
// rowptr[nrows+1], colind/valT/valA[9*nrows] (entries with value 0 in both matrices are dropped)
int scftb2d_export_csr(scftb2d_engine *e, int *rowptr, int *colind, double *valT, double *valA) {
  if (!e || !rowptr || !colind || !valT || !valA) return fail(SCFTB_ERR_ARG, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  const size_t nv = (size_t)e->nslices * SLOTS * 32;
  std::vector<int> col(nv);
  std::vector<double> vt(nv), va(nv);
  CK(cudaMemcpy(col.data(), e->M.S.col, sizeof(int) * nv, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vt.data(), e->M.S.valT, sizeof(double) * nv, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(va.data(), e->M.S.valA, sizeof(double) * nv, cudaMemcpyDeviceToHost));
  int nnz = 0;
  for (int i = 0; i < e->nrows; i++) {
    rowptr[i] = nnz;
    for (int k = 0; k < SLOTS; k++) {
      size_t s = (size_t)(i >> 5) * (SLOTS * 32) + k * 32 + (i & 31);
      if (vt[s] == 0.0 && va[s] == 0.0) continue;
      colind[nnz] = col[s] - e->halo + e->ix0 * e->nyp;
      valT[nnz] = vt[s]; valA[nnz] = va[s];
      nnz++;
    }
  }
  rowptr[e->nrows] = nnz;
  return SCFTB_OK;
}

}  // extern "C"
