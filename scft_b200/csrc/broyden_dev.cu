// broyden_dev.cu — device-resident Broyden (broydn.c:44-292 with fdjac.c, qrdcmp.c, qrupdt.c,
// rotate.c, rsolv.c, lnsrch.c): the Jacobian, its QR factors, the rank-one Givens updates, the
// triangular solves and the line-search vectors never leave the GPU; the host only steers the
// iteration with a handful of scalars per step (|f|max, 1/2 f.f, slope, step tests).
//
//   fdjac      n perturbed residual evaluations = ONE batched march launch (fdjac.c:18-34)
//   qrdcmp     Householder QR, one cooperative kernel: a warp per 32 columns, grid barrier per reflector
//   Q^T        every column of Q^T evolves independently under the reflectors (broydn.c:129-149): no barriers
//   qrupdt     Givens sweeps (qrupdt.c:5-24, rotate.c:5-34) in one CTA, a thread per column, the running
//              row kept in registers
//   rsolv      back substitution, one CTA, block-wide dot product per row (rsolv.c:3-13)
// The iteration logic, tolerances and return conventions are those of scftb_broydn (solvers.cu),
// which reproduces the reference's broydn bit for bit; sums here are reduced in parallel, so the
// two agree to rounding, not bitwise.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.h"

namespace cg = cooperative_groups;
using namespace scftb;

namespace {

constexpr int BT = 256;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ double dsign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// block reductions (blockDim.x = BT)
__device__ double block_sum(double v) {
  __shared__ double sh[BT / 32];
  __shared__ double tot;
  v = wsum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < BT / 32; i++) s += sh[i]; tot = s; }
  __syncthreads();
  double r = tot;
  __syncthreads();
  return r;
}
__device__ double block_max(double v) {
  __shared__ double sh[BT / 32];
  __shared__ double tot;
  v = wmax(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < BT / 32; i++) s = fmax(s, sh[i]); tot = s; }
  __syncthreads();
  double r = tot;
  __syncthreads();
  return r;
}

// ---- fdjac.c:18-34 ---------------------------------------------------------------------------------
__global__ void jac_inputs_kernel(int n, const double *x, double *xb, double *hs) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int j = idx / n, i = idx - (size_t)j * n;
  double v = x[i];
  if (i == j) {
    const double EPS = 1.0e-7;
    double h = EPS * v;
    if (fabs(h) < EPS) h = dsign(EPS, v);
    const double xp = v + h;
    hs[j] = xp - v;
    v = xp;
  }
  xb[idx] = v;
}
__global__ void jac_form_kernel(int n, const double *fb, const double *fvec, const double *hs, double *r) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int i = idx / n, j = idx - (size_t)i * n;
  r[idx] = (fb[(size_t)j * n + i] - fvec[i]) / hs[j];
}

// ---- qrdcmp.c:5-33: one warp per 32 columns, every warp recomputes the reflector from column k ---------
__global__ void __launch_bounds__(32) qrdcmp_kernel(int n, double *a, double *c, double *d, int *sing) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double vk[];   // normalised column k, rows k..n-1
  const int lane = threadIdx.x, j = blockIdx.x * 32 + lane;
  if (blockIdx.x == 0 && lane == 0) *sing = 0;
  for (int k = 0; k < n - 1; k++) {
    double sc = 0.0;
    for (int i = k + lane; i < n; i += 32) sc = fmax(sc, fabs(a[(size_t)i * n + k]));
    sc = wmax(sc);
    if (sc == 0.0) {
      if (blockIdx.x == 0 && lane == 0) { *sing = 1; c[k] = 0.0; d[k] = 0.0; }
    } else {
      double ss = 0.0;
      for (int i = k + lane; i < n; i += 32) { double v = a[(size_t)i * n + k] / sc; vk[i - k] = v; ss = fma(v, v, ss); }
      ss = wsum(ss);
      __syncwarp();
      const double sigma = dsign(sqrt(ss), vk[0]);
      const double akk = vk[0] + sigma, ck = sigma * akk;
      __syncwarp();
      if (lane == 0) vk[0] = akk;
      __syncwarp();
      if (j > k && j < n) {   // apply the reflector to the own column
        double sum = 0.0;
        for (int i = k; i < n; i++) sum = fma(vk[i - k], a[(size_t)i * n + j], sum);
        const double tau = sum / ck;
        for (int i = k; i < n; i++) a[(size_t)i * n + j] = fma(-tau, vk[i - k], a[(size_t)i * n + j]);
      }
      if (k / 32 == blockIdx.x && lane == 0) { c[k] = ck; d[k] = -sc * sigma; }
    }
    // every warp reads column k at the top of this step, so its owner may overwrite it with the reflector only
    // after the grid barrier (nobody touches column k any more in the later steps)
    grid.sync();
    if (sc != 0.0 && k / 32 == blockIdx.x) {
      for (int i = k + lane; i < n; i += 32) a[(size_t)i * n + k] = vk[i - k];
      __syncwarp();
    }
  }
  if (blockIdx.x == 0 && lane == 0) {
    d[n - 1] = a[(size_t)(n - 1) * n + n - 1];
    if (d[n - 1] == 0.0) *sing = 1;
  }
}

// ---- Q^T = H_{n-2} ... H_0 applied to the identity, column by column (broydn.c:129-149) -----------------
__global__ void __launch_bounds__(32) form_qt_kernel(int n, const double *r, const double *c, double *qt) {
  extern __shared__ double vk[];
  const int lane = threadIdx.x, j = blockIdx.x * 32 + lane;
  for (int i = 0; i < n; i++)
    if (j < n) qt[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int k = 0; k < n - 1; k++) {
    const double ck = c[k];
    if (ck == 0.0) continue;
    __syncwarp();
    for (int i = k + lane; i < n; i += 32) vk[i - k] = r[(size_t)i * n + k];
    __syncwarp();
    if (j < n) {
      // column j of Q^T is still e_j below row max(j,k): rows i > j of the column are zero until a reflector
      // with k <= j touches them, so the sum starts at i = k (all rows may be non-zero once k <= j)
      double sum = 0.0;
      for (int i = k; i < n; i++) sum = fma(vk[i - k], qt[(size_t)i * n + j], sum);
      sum /= ck;
      for (int i = k; i < n; i++) qt[(size_t)i * n + j] = fma(-sum, vk[i - k], qt[(size_t)i * n + j]);
    }
  }
}
__global__ void finish_r_kernel(int n, double *r, const double *d) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int i = idx / n, j = idx - (size_t)i * n;
  if (j < i) r[idx] = 0.0;
  else if (j == i) r[idx] = d[i];
}

// ---- dense helpers: a warp per row (row-wise dot) or a thread per column (column-wise dot) ---------------
// mode 0: y_i = sum_j A(i,j) x_j over j in [lo_i, n)   with lo_i = (tri ? i : 0);   y scaled by `scale`
__global__ void rowdot_kernel(int n, const double *A, const double *x, double *y, int tri, double scale) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  double s = 0.0;
  for (int j = (tri ? w : 0) + lane; j < n; j += 32) s = fma(A[(size_t)w * n + j], x[j], s);
  s = wsum(s);
  if (lane == 0) y[w] = scale * s;
}
// y_i = sum_j A(j,i) x_j over j in [0, hi_i)   with hi_i = (tri ? i+1 : n)
__global__ void coldot_kernel(int n, const double *A, const double *x, double *y, int tri, double scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  const int hi = tri ? i + 1 : n;
  for (int j = 0; j < hi; j++) s = fma(A[(size_t)j * n + i], x[j], s);
  y[i] = scale * s;
}

// ---- rsolv.c:3-13 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) rsolv_kernel(int n, const double *a, const double *d, double *b) {
  extern __shared__ double bs[];   // solution so far
  for (int i = threadIdx.x; i < n; i += BT) bs[i] = b[i];
  __syncthreads();
  for (int i = n - 1; i >= 0; i--) {
    double s = 0.0;
    for (int j = i + 1 + threadIdx.x; j < n; j += BT) s = fma(a[(size_t)i * n + j], bs[j], s);
    s = block_sum(s);
    if (threadIdx.x == 0) bs[i] = (bs[i] - s) / d[i];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += BT) b[i] = bs[i];
}

// ---- qrupdt.c:5-24 with rotate.c:5-34; one CTA, thread-strided columns ----------------------------------
__device__ __forceinline__ void givens(double a, double b, double &c, double &s) {
  if (a == 0.0) { c = 0.0; s = (b >= 0.0 ? 1.0 : -1.0); }
  else if (fabs(a) > fabs(b)) { double f = b / a; c = dsign(1.0 / sqrt(1.0 + f * f), a); s = f * c; }
  else { double f = a / b; s = dsign(1.0 / sqrt(1.0 + f * f), b); c = f * s; }
}
__global__ void __launch_bounds__(1024) qrupdt_kernel(int n, double *r, double *qt, const double *u_in, const double *v, double *d,
                                                      int *rsing) {
  extern __shared__ double u[];   // [n] + 2 scratch
  double *ab = u + n;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) u[i] = u_in[i];
  __syncthreads();
  int k = n - 1;
  while (k >= 0 && u[k] == 0.0) k--;
  if (k < 0) k = 0;
  // first sweep: rotations i = k-1 .. 0 zero u below its first entry, turning R into upper Hessenberg
  for (int i = k - 1; i >= 0; i--) {
    double c, s;
    givens(u[i], -u[i + 1], c, s);
    for (int j = tid; j < n; j += nt) {
      if (j >= i) { double y = r[(size_t)i * n + j], w = r[(size_t)(i + 1) * n + j]; r[(size_t)i * n + j] = c * y - s * w; r[(size_t)(i + 1) * n + j] = s * y + c * w; }
      double y = qt[(size_t)i * n + j], w = qt[(size_t)(i + 1) * n + j];
      qt[(size_t)i * n + j] = c * y - s * w; qt[(size_t)(i + 1) * n + j] = s * y + c * w;
    }
    __syncthreads();
    if (tid == 0) {
      if (u[i] == 0.0) u[i] = fabs(u[i + 1]);
      else if (fabs(u[i]) > fabs(u[i + 1])) { double q = u[i + 1] / u[i]; u[i] = fabs(u[i]) * sqrt(1.0 + q * q); }
      else { double q = u[i] / u[i + 1]; u[i] = fabs(u[i + 1]) * sqrt(1.0 + q * q); }
    }
    __syncthreads();
  }
  for (int j = tid; j < n; j += nt) r[j] += u[0] * v[j];
  __syncthreads();
  // second sweep: rotations i = 0 .. k-1 restore the triangle
  for (int i = 0; i < k; i++) {
    if (tid == 0) { ab[0] = r[(size_t)i * n + i]; ab[1] = -r[(size_t)(i + 1) * n + i]; }
    __syncthreads();
    double c, s;
    givens(ab[0], ab[1], c, s);
    for (int j = tid; j < n; j += nt) {
      if (j >= i) { double y = r[(size_t)i * n + j], w = r[(size_t)(i + 1) * n + j]; r[(size_t)i * n + j] = c * y - s * w; r[(size_t)(i + 1) * n + j] = s * y + c * w; }
      double y = qt[(size_t)i * n + j], w = qt[(size_t)(i + 1) * n + j];
      qt[(size_t)i * n + j] = c * y - s * w; qt[(size_t)(i + 1) * n + j] = s * y + c * w;
    }
    __syncthreads();
  }
  if (tid == 0) *rsing = 0;
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    double dv = r[(size_t)i * n + i];
    if (dv == 0.0) *rsing = 1;
    d[i] = dv;
  }
}

// ---- small vector kernels ----------------------------------------------------------------------------
// out[0] = 1/2 f.f, out[1] = max|f|, out[2] = any NaN
__global__ void __launch_bounds__(BT) fnorm_kernel(int n, const double *f, double *out) {
  double s = 0.0, m = 0.0, bad = 0.0;
  for (int i = threadIdx.x; i < n; i += BT) { double v = f[i]; s = fma(v, v, s); if (fabs(v) > m) m = fabs(v); if (isnan(v)) bad = 1.0; }
  s = block_sum(s); m = block_max(m); bad = block_max(bad);
  if (threadIdx.x == 0) { out[0] = 0.5 * s; out[1] = m; out[2] = bad; }
}
// out[0] = x.x, out[1] = p.p, out[2] = g.p, out[3] = max |p_i| / max(|xold_i|,1), out[4] = max |g_i| max(|x_i|,1),
// out[5] = max |x_i - xold_i| / max(|x_i|,1)
__global__ void __launch_bounds__(BT) stats_kernel(int n, const double *x, const double *xold, const double *p, const double *g,
                                                   double *out) {
  double a = 0, b = 0, c = 0, t1 = 0, t2 = 0, t3 = 0;
  for (int i = threadIdx.x; i < n; i += BT) {
    a = fma(x[i], x[i], a); b = fma(p[i], p[i], b); c = fma(g[i], p[i], c);
    t1 = fmax(t1, fabs(p[i]) / fmax(fabs(xold[i]), 1.0));
    t2 = fmax(t2, fabs(g[i]) * fmax(fabs(x[i]), 1.0));
    t3 = fmax(t3, fabs(x[i] - xold[i]) / fmax(fabs(x[i]), 1.0));
  }
  a = block_sum(a); b = block_sum(b); c = block_sum(c); t1 = block_max(t1); t2 = block_max(t2); t3 = block_max(t3);
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = c; out[3] = t1; out[4] = t2; out[5] = t3; }
}
__global__ void axpy_kernel(int n, const double *xold, const double *p, double alam, double *x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = xold[i] + alam * p[i];
}
__global__ void scale_kernel(int n, double *p, double f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= f;
}
// s = x - xold ; returns nothing
__global__ void diff_kernel(int n, const double *x, const double *xold, double *s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) s[i] = x[i] - xold[i];
}
// w_i = fvec_i - fvcold_i - w_i (w holds Q t on entry); noise filter of broydn.c:174-177; out[0] = skip flag; out[1] = s.s
__global__ void __launch_bounds__(BT) wfilter_kernel(int n, const double *fvec, const double *fvcold, double *w, const double *s,
                                                     double *out) {
  const double EPS = 1e-14;
  double any = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < n; i += BT) {
    double wi = fvec[i] - fvcold[i] - w[i];
    if (fabs(wi) >= EPS * (fabs(fvec[i]) + fabs(fvcold[i]))) any = 1.0; else wi = 0.0;
    w[i] = wi;
    den = fma(s[i], s[i], den);
  }
  any = block_max(any); den = block_sum(den);
  if (threadIdx.x == 0) { out[0] = (any == 0.0) ? 1.0 : 0.0; out[1] = den; }
}

}  // namespace

struct BroydenDev {   // plays the role of the caller-owned globals qt, r, d (broydn.c:22-28): kept per engine across calls for jc
  int n = 0;
  double *r = nullptr, *qt = nullptr, *xb = nullptr, *fb = nullptr, *vec = nullptr, *scal = nullptr;
  int *flags = nullptr;
  void release() {
    for (void *p : {(void *)r, (void *)qt, (void *)xb, (void *)fb, (void *)vec, (void *)scal, (void *)flags})
      if (p) cudaFree(p);
    r = qt = xb = fb = vec = scal = nullptr; flags = nullptr; n = 0;
  }
};
static void free_broyden_state(void *p) {
  BroydenDev *b = (BroydenDev *)p;
  b->release();
  delete b;
}

extern "C" int scftb_broydn_device(scftb_engine *e, double *x_host, int *check, double *err, int *jc) {
  return scftb_broydn_device_ex(e, x_host, check, err, jc, 0);
}

extern "C" int scftb_broydn_device_ex(scftb_engine *e, double *x_host, int *check, double *err, int *jc, int flags) {
  if (!e || !x_host || !check || !err || !jc) return fail(SCFTB_ERR_ARG, "broydn_device: bad argument");
  const bool keep_trial = (flags & SCFTB_BROYDN_KEEP_TRIAL) != 0;
  const int n = e->ni;
  const int MAXITS = 400;
  const double TOLX = 1e-14, STPMX = 100.0, TOLF = *err, TOLMIN = TOLF;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = e->stream;
  int rc = upload_params(e);
  if (rc) return rc;
  if (!e->solver_state) { e->solver_state = new BroydenDev(); e->solver_state_free = free_broyden_state; }
  BroydenDev &g_bd = *(BroydenDev *)e->solver_state;
  if (g_bd.n != n) {
    g_bd.release();
    const size_t nn = (size_t)n * n;
    CK(cudaMalloc(&g_bd.r, 8 * nn)); CK(cudaMalloc(&g_bd.qt, 8 * nn)); CK(cudaMalloc(&g_bd.xb, 8 * nn)); CK(cudaMalloc(&g_bd.fb, 8 * nn));
    CK(cudaMalloc(&g_bd.vec, 8 * (size_t)n * 12)); CK(cudaMalloc(&g_bd.scal, 8 * 16)); CK(cudaMalloc(&g_bd.flags, sizeof(int) * 4));
    g_bd.n = n;
    *jc = 0;
  }
  double *r = g_bd.r, *qt = g_bd.qt, *V = g_bd.vec;
  double *x = V, *xold = V + n, *fvec = V + 2 * n, *fvcold = V + 3 * n, *g = V + 4 * n, *p = V + 5 * n, *s = V + 6 * n, *t = V + 7 * n,
         *w = V + 8 * n, *c = V + 9 * n, *d = V + 10 * n, *hs = V + 11 * n;
  double hsc[8];
  int hfl[4];
  const int GB = (n + 255) / 256, GNN = (int)(((size_t)n * n + 255) / 256), GW = (n * 32 + 255) / 256, GC = (n + 31) / 32;
  const size_t smem_col = sizeof(double) * (n + 2);
  auto eval = [&](double &fval, double &emax) -> int {   // fminbrd (broydn.c:30-42) + |f|max
    int q = launch_march(e, 1, x, n, fvec, n, nullptr, st);
    if (q) return q;
    fnorm_kernel<<<1, BT, 0, st>>>(n, fvec, g_bd.scal);
    g_launches++;
    CK(cudaMemcpyAsync(hsc, g_bd.scal, 8 * 3, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    fval = hsc[0]; emax = hsc[1];
    return hsc[2] != 0.0 ? SCFTB_ERR_NAN : SCFTB_OK;
  };
  auto stats = [&]() -> int {
    stats_kernel<<<1, BT, 0, st>>>(n, x, xold, p, g, g_bd.scal + 8);
    g_launches++;
    CK(cudaMemcpyAsync(hsc, g_bd.scal + 8, 8 * 6, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SCFTB_OK;
  };
  auto finish = [&](int status) -> int {
    CK(cudaMemcpyAsync(x_host, x, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return status;
  };
  if ((int)e->cfg.max_batch < 1) return fail(SCFTB_ERR_ARG, "engine without capacity");
  CK(cudaMemcpyAsync(x, x_host, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(xold, x_host, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  double f, emax;
  rc = eval(f, emax);
  if (rc) { *check = 1; return finish(rc == SCFTB_ERR_NAN ? SCFTB_OK : rc); }
  *err = emax;
  if (emax < TOLF) { *check = 0; return finish(SCFTB_OK); }
  if ((rc = stats())) return rc;
  const double stpmax = STPMX * std::max(std::sqrt(hsc[0]), (double)n);
  int restrt = (*jc == 0) ? 1 : 0;
  *check = 1;

  for (int its = 1; its <= MAXITS; its++) {
    if (restrt) {
      jac_inputs_kernel<<<GNN, 256, 0, st>>>(n, x, g_bd.xb, hs);
      g_launches++;
      const int B = e->cfg.max_batch;
      for (int j0 = 0; j0 < n; j0 += B) {
        const int nb = std::min(B, n - j0);
        // pshare: all columns with problem 0's (tau, L, mesh); the other slots' parameters and outputs are not involved
        if ((rc = launch_march(e, nb, g_bd.xb + (size_t)j0 * n, n, g_bd.fb + (size_t)j0 * n, n, nullptr, st, 0, true))) return rc;
      }
      jac_form_kernel<<<GNN, 256, 0, st>>>(n, g_bd.fb, fvec, hs, r);
      {
        int nn_ = n;
        void *args[] = {&nn_, &r, &c, &d, &g_bd.flags};
        CK(cudaLaunchCooperativeKernel((void *)qrdcmp_kernel, dim3(GC), dim3(32), args, smem_col, st));
      }
      form_qt_kernel<<<GC, 32, smem_col, st>>>(n, r, c, qt);
      finish_r_kernel<<<GNN, 256, 0, st>>>(n, r, d);
      g_launches += 4;
      CK(cudaMemcpyAsync(hfl, g_bd.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (hfl[0]) { *check = 1; finish(SCFTB_OK); return fail(SCFTB_ERR_NOCONV, "singular Jacobian in broydn"); }
      *jc = 2;
    } else if (its > 1) {   // Broyden update (broydn.c:158-199)
      diff_kernel<<<GB, 256, 0, st>>>(n, x, xold, s);
      rowdot_kernel<<<GW, 256, 0, st>>>(n, r, s, t, 1, 1.0);           // t = R s
      coldot_kernel<<<GB, 256, 0, st>>>(n, qt, t, w, 0, 1.0);           // w = Q t  (sum_j qt(j,i) t_j)
      wfilter_kernel<<<1, BT, 0, st>>>(n, fvec, fvcold, w, s, g_bd.scal);
      g_launches += 4;
      CK(cudaMemcpyAsync(hsc, g_bd.scal, 8 * 2, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (hsc[0] == 0.0) {   // not skipped
        rowdot_kernel<<<GW, 256, 0, st>>>(n, qt, w, t, 0, 1.0);         // t = Q^T w
        scale_kernel<<<GB, 256, 0, st>>>(n, s, 1.0 / hsc[1]);
        qrupdt_kernel<<<1, 1024, smem_col, st>>>(n, r, qt, t, s, d, g_bd.flags + 1);
        g_launches += 3;
        CK(cudaMemcpyAsync(hfl, g_bd.flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (hfl[0]) { *check = 1; finish(SCFTB_OK); return fail(SCFTB_ERR_NOCONV, "r singular in broydn"); }
      }
    }
    rowdot_kernel<<<GW, 256, 0, st>>>(n, qt, fvec, p, 0, -1.0);         // p = -Q^T f
    coldot_kernel<<<GB, 256, 0, st>>>(n, r, p, g, 1, -1.0);             // g = -R^T p  (sum_{j<=i} r(j,i) p_j)
    CK(cudaMemcpyAsync(xold, x, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(fvcold, fvec, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    const double fold = f;
    rsolv_kernel<<<1, BT, sizeof(double) * n, st>>>(n, r, d, p);
    g_launches += 3;
    // ---- lnsrch.c:4-60
    if ((rc = stats())) return rc;
    double pn = std::sqrt(hsc[1]), slope = hsc[2], test = hsc[3];
    if (pn > stpmax) { scale_kernel<<<GB, 256, 0, st>>>(n, p, stpmax / pn); slope *= stpmax / pn; test *= stpmax / pn; g_launches++; }
    if (slope >= 0.0) { *check = 1; finish(SCFTB_OK); return fail(SCFTB_ERR_NOCONV, "Roundoff problem in lnsrch"); }
    const double ALF = 1.0e-4, LTOLX = 1.0e-7, alamin = LTOLX / test;
    double alam = 1.0, alam2 = 0.0, f2 = 0.0, tmplam;
    *check = 0;
    for (;;) {
      axpy_kernel<<<GB, 256, 0, st>>>(n, xold, p, alam, x);
      g_launches++;
      rc = eval(f, emax);
      if (rc == SCFTB_ERR_NAN) { f = INFINITY; emax = INFINITY; }   // a NaN trial behaves like a rejected step
      else if (rc) return rc;
      if (alam < alamin) {
        if (keep_trial && emax < TOLF) break;   // the trial itself is a solution: return it, not the previous iterate
        CK(cudaMemcpyAsync(x, xold, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st)); *check = 1; break;
      }
      else if (f <= fold + ALF * alam * slope) break;
      else {
        if (alam == 1.0) tmplam = -slope / (2.0 * (f - fold - slope));
        else {
          double rhs1 = f - fold - alam * slope, rhs2 = f2 - fold - alam2 * slope;
          double a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
          double b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
          if (a == 0.0) tmplam = -slope / (2.0 * b);
          else {
            double disc = b * b - 3.0 * a * slope;
            if (disc < 0.0) tmplam = 0.5 * alam;
            else if (b <= 0.0) tmplam = (-b + std::sqrt(disc)) / (3.0 * a);
            else tmplam = -slope / (b + std::sqrt(disc));
          }
          if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
        }
        if (!(tmplam == tmplam)) tmplam = 0.1 * alam;   // f = inf from a NaN trial
      }
      alam2 = alam; f2 = f;
      alam = std::max(tmplam, 0.1 * alam);
    }
    *err = emax;
    if (emax < TOLF) { *check = 0; *jc = 1; return finish(SCFTB_OK); }
    if (*check) {
      if (restrt) return finish(SCFTB_OK);
      if ((rc = stats())) return rc;
      const double den = std::max(f, 0.5 * n);
      if (hsc[4] / den < TOLMIN) { *check = 0; *jc = 1; return finish(SCFTB_OK); }
      restrt = 1;
    } else {
      restrt = 0;
      if ((rc = stats())) return rc;
      if (hsc[5] < TOLX) { *jc = 1; return finish(SCFTB_OK); }
    }
  }
  *check = 1;
  return finish(SCFTB_OK);
}
