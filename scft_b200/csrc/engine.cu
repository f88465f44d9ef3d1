// engine.cu — C-ABI implementation of the SCFT propagator engine (see include/scft_b200.h).
// Host side: owns the device buffers of up to max_batch independent problems, computes the
// per-mesh constants (phi_0, quadrature weights) once, and launches the march kernel.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "march1d.cuh"
#include "march_irk4.cuh"
#include "march1d_tmem.cuh"
#include "march_irk4_tm.cuh"

namespace scftb {

thread_local std::string g_last_error;
std::atomic<long> g_launches{0};

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

// ---------------------------------------------------------------------------------------------
// Romberg quadrature as a weight vector.  romint.c:21-57 builds trapezoid sums on 1,2,4,...,m
// intervals and extrapolates the last K=5 of them to h^2 -> 0 with Neville's scheme
// (polint.c:5-42).  Both are linear in f, so the result is sum_j w_j f_j with
// w = sum_k lagrange_k(0) * (trapezoid weights of level k), h_k = 4^-k.
// ---------------------------------------------------------------------------------------------
int romberg_weights(int m, double hh, std::vector<double> &w) {
  int levels = 0;
  while ((1 << levels) < m) levels++;
  if ((1 << levels) != m || levels < 4) return 1;  // romint.c:29-33 needs m = 2^k >= 16
  const int K = 5, M = levels + 1;                 // trapezoid levels 1..M, level l has 2^(l-1) intervals
  double hk[K], ck[K];
  for (int k = 0; k < K; k++) hk[k] = std::pow(0.25, (double)(M - K + k));  // h[l] = 4^-(l-1)
  for (int k = 0; k < K; k++) {  // Lagrange basis at 0
    double c = 1.0;
    for (int j = 0; j < K; j++)
      if (j != k) c *= (0.0 - hk[j]) / (hk[k] - hk[j]);
    ck[k] = c;
  }
  w.assign(m + 1, 0.0);
  for (int k = 0; k < K; k++) {
    int level = M - K + 1 + k;
    int stride = m >> (level - 1);
    double hl = hh * stride;
    for (int i = 0; i <= m; i += stride) w[i] += ck[k] * hl * ((i == 0 || i == m) ? 0.5 : 1.0);
  }
  return 0;
}

void trapezoid_weights(int m, double hh, std::vector<double> &w) {  // simple_FEM_1D_transient.m:120-124
  w.assign(m + 1, hh);
  w[0] = w[m] = 0.5 * hh;
}

// phi_0 on the nodes (scft.cc:188-215): ((e^u-1)/(e^u+1))^2, u = 4 tau x/(tau^2-x^2), x <= tau,
// mirrored about the film centre, NaN -> 1.
static double f0_point(double x, double tau) {
  double e = std::exp(4 * tau * x / (tau * tau - x * x));
  double v = std::pow(e - 1, 2) / std::pow(e + 1, 2);
  return std::isnan(v) ? 1.0 : v;
}
void f0_given(int N, const double *x, double tau, double *f0) {
  for (int i = 0; i < N; i++) f0[i] = 1.0;
  for (int i = 0; i < N; i++) {
    if (x[i] <= tau) {
      f0[i] = f0_point(x[i], tau);
      f0[N - i - 1] = f0[i];
    } else
      break;
  }
}

// ---------------------------------------------------------------------------------------------
// natural-spline wall values on a NON-uniform mesh (scft.cc:452-490, spline_chen.c:12-106 with
// m = 0): the reference solves the tridiagonal system densely with gaussj; here one thread per
// problem runs Thomas over the N-2 knots, its two sweep arrays interleaved across the problems of the
// launch (element i of problem p at [i*nprob + p]) so that a warp's scratch accesses are contiguous.
// Uniform meshes never need this (see eta_node()).
// ---------------------------------------------------------------------------------------------
__global__ void spline_bnd_kernel(int nprob, int N, const double *x, const double *eta_mid, long long eta_stride,
                                  double *scratch, double *eta_bnd, int pshare) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  const int Nx = N - 2;
  x += (size_t)(pshare ? 0 : p) * N;         // mesh of this problem (pshare: every problem on the mesh of slot 0)
  const double *xk = x + 1;                  // knots = interior nodes
  const double *y = eta_mid + (size_t)p * eta_stride;
  const size_t S = nprob;
  double *cp = scratch + p, *dp = scratch + (size_t)Nx * S + p;
  // rows i=1..Nx-2: (x_i-x_{i-1})/6, (x_{i+1}-x_{i-1})/3, (x_{i+1}-x_i)/6 ; rows 0, Nx-1: M = 0
  double cprev = 0.0, dprev = 0.0;
  for (int i = 1; i < Nx - 1; i++) {
    double lo = (xk[i] - xk[i - 1]) / 6., di = (xk[i + 1] - xk[i - 1]) / 3., up = (xk[i + 1] - xk[i]) / 6.;
    double rhs = (y[i + 1] - y[i]) / (xk[i + 1] - xk[i]) - (y[i] - y[i - 1]) / (xk[i] - xk[i - 1]);
    double den = di - lo * cprev;
    cprev = up / den;
    dprev = (rhs - lo * dprev) / den;
    cp[i * S] = cprev; dp[i * S] = dprev;
  }
  double Mn = 0.0, M1 = 0.0, Mlast1 = 0.0;
  for (int i = Nx - 2; i >= 1; i--) {
    Mn = dp[i * S] - cp[i * S] * Mn;
    if (i == Nx - 2) Mlast1 = Mn;
    if (i == 1) M1 = Mn;
  }
  const double *xf = x;
  {  // left wall: klo=0, khi=1
    double h = xk[1] - xk[0], a = (xk[1] - xf[0]) / h, b = (xf[0] - xk[0]) / h;
    eta_bnd[2 * p] = a * y[0] + b * y[1] + ((a * a * a - a) * 0.0 + (b * b * b - b) * M1) * (h * h) / 6.0;
  }
  {  // right wall: klo=Nx-2, khi=Nx-1
    double h = xk[Nx - 1] - xk[Nx - 2], a = (xk[Nx - 1] - xf[N - 1]) / h, b = (xf[N - 1] - xk[Nx - 2]) / h;
    eta_bnd[2 * p + 1] =
        a * y[Nx - 2] + b * y[Nx - 1] + ((a * a * a - a) * Mlast1 + (b * b * b - b) * 0.0) * (h * h) / 6.0;
  }
}

void spline_bnd_launch(int nprob, int N, const double *x, const double *eta_mid, long long eta_stride, double *scratch,
                       double *eta_bnd, cudaStream_t st, int pshare) {
  spline_bnd_kernel<<<(nprob + 63) / 64, 64, 0, st>>>(nprob, N, x, eta_mid, eta_stride, scratch, eta_bnd, pshare);
  g_launches++;
}

// ---------------------------------------------------------------------------------------------
template <int C, int T, int MINB>
static march_fn pick(bool uni, bool odd) {
  if (odd) return uni ? (march_fn)march_ie_kernel<C, T, true, MINB, true> : (march_fn)march_ie_kernel<C, T, false, MINB, true>;
  return uni ? (march_fn)march_ie_kernel<C, T, true, MINB, false> : (march_fn)march_ie_kernel<C, T, false, MINB, false>;
}

// nodes per thread C and threads per problem T for ni interior nodes.  SCFTB_FORCE_C=4 selects the
// 256-thread / 4-nodes-per-thread variant where it applies (tuning experiments).
int choose_kernel(int ni, bool uni, KernelChoice &kc, bool odd) {
  int C = 1;
  while (C < 16 && (ni + C - 1) / C > 128) C *= 2;
  // small meshes: a contour step costs about the same ~1100 cycles per CTA whatever C is (it is the dependent chain through
  // the three levels), so up to 8 nodes per thread and ONE warp per problem where that fits (no level 3, no barrier) multiplies
  // the problems in flight per SM: the coarse levels of a sweep's continuation run 2-4x faster (SCFTB_NARROW_CHUNKS=1: old rule)
  const char *narrow = getenv("SCFTB_NARROW_CHUNKS");
  if (!(narrow && atoi(narrow))) {
    if (ni > 32 && ni <= 64) C = 2;
    else if (ni > 64 && ni <= 128) C = 4;
    else if (ni > 128 && ni <= 512) C = 8;
  }
  const char *force = getenv("SCFTB_FORCE_C");
  if (force && atoi(force) == 4 && C == 8 && ni > 512) C = 4;
  int need = (ni + C - 1) / C;
  int T = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 256));
  if (need > 256) return 1;
  kc.fn = nullptr;
  if (C == 2 && T == 32) kc.fn = pick<2, 32, 16>(uni, odd);
  if (C == 4 && T == 32) kc.fn = pick<4, 32, 16>(uni, odd);
  if (C == 8 && T == 32) kc.fn = pick<8, 32, 12>(uni, odd);
  if (C == 8 && T == 64) kc.fn = pick<8, 64, 6>(uni, odd);
  if (C == 1 && T == 32) kc.fn = pick<1, 32, 8>(uni, odd);
  if (C == 1 && T == 64) kc.fn = pick<1, 64, 6>(uni, odd);
  if (C == 1 && T == 128) kc.fn = pick<1, 128, 4>(uni, odd);
  if (C == 2 && T == 128) kc.fn = pick<2, 128, 4>(uni, odd);
  if (C == 4 && T == 128) kc.fn = pick<4, 128, 4>(uni, odd);
  if (C == 4 && T == 256) kc.fn = pick<4, 256, 2>(uni, odd);
  if (C == 8 && T == 128) kc.fn = pick<8, 128, 3>(uni, odd);
  if (C == 16 && T == 128) kc.fn = pick<16, 128, 1>(uni, odd);
  if (C == 16 && T == 256) kc.fn = pick<16, 256, 1>(uni, odd);
  if (!kc.fn) return 1;
  kc.C = C; kc.T = T; kc.dyn_smem = 0; kc.max_occ = 0;
  // the benchmarked shape (513..1024 unknowns on a uniform mesh, even step count): coefficients in tensor memory,
  // <= 128 registers, four CTAs per SM.  SCFTB_NO_TMEM=1 keeps the register-resident kernel (A/B measurements).
  const char *no_tm = getenv("SCFTB_NO_TMEM");
  if (C == 8 && T == 128 && uni && !odd && !(no_tm && atoi(no_tm))) {
    kc.fn = (march_fn)march_tm_kernel;
    kc.dyn_smem = 12 * 1024;   // static + dynamic > 228 KB / 5: a fifth CTA (no tensor-memory columns left) never lands
    kc.max_occ = 4;
  }
  return 0;
}

// IRK4 (complex solve): fewer nodes per thread, twice the coefficient storage
template <int C, int T>
static march_fn pick4(bool uni) {
  return uni ? (march_fn)march_irk4_kernel<C, T, true> : (march_fn)march_irk4_kernel<C, T, false>;
}

int choose_kernel_irk4(int ni, bool uni, KernelChoice &kc, int max_batch) {
  int C = 1;
  while (C < 8 && (ni + C - 1) / C > 256) C *= 2;
  int need = (ni + C - 1) / C;
  if (need > 256) return 1;
  int T = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 256));
  kc.fn = nullptr;
  if (C == 1 && T == 32) kc.fn = pick4<1, 32>(uni);
  if (C == 1 && T == 64) kc.fn = pick4<1, 64>(uni);
  if (C == 1 && T == 128) kc.fn = pick4<1, 128>(uni);
  if (C == 1 && T == 256) kc.fn = pick4<1, 256>(uni);
  if (C == 2 && T == 256) kc.fn = pick4<2, 256>(uni);
  if (C == 4 && T == 256) kc.fn = pick4<4, 256>(uni);
  if (C == 8 && T == 256) kc.fn = pick4<8, 256>(uni);
  if (!kc.fn) return 1;
  kc.C = C; kc.T = T; kc.dyn_smem = 0; kc.max_occ = 0;
  // 513..1024 unknowns on a uniform mesh: complex coefficients in tensor memory (march_irk4_tm.cuh), two CTAs per SM — a
  // throughput kernel (+76 % on sweeps); the tcgen05.ld waits sit on a lone CTA's critical path (the C = 4 version measured
  // 2253 vs 1925 cycles per step), so engines that can never fill more than one CTA per SM (max_batch <= 148: single
  // problems, small Jacobian batches) keep the register-resident kernel
  const char *no_tm = getenv("SCFTB_NO_TMEM");
  if (C == 4 && T == 256 && uni && max_batch > 148 && !(no_tm && atoi(no_tm))) {
    kc.fn = (march_fn)march_irk4_tm_kernel;
    kc.C = 8; kc.T = 128;   // 8 nodes per thread: one separator per 8 nodes
    kc.max_occ = 2;
  }
  return 0;
}

static int choose_any(int scheme, int ni, bool uni, KernelChoice &kc, int nsteps, int max_batch) {
  return scheme == SCFTB_IRK4_CONSISTENT ? choose_kernel_irk4(ni, uni, kc, max_batch) : choose_kernel(ni, uni, kc, (nsteps & 1) != 0);
}

}  // namespace scftb

using namespace scftb;

namespace scftb {
int upload_params(scftb_engine *e) {
  if (!e->params_dirty) return SCFTB_OK;
  const int N = e->cfg.N, B = e->cfg.max_batch;
  CK(cudaMemcpyAsync(e->d_f0, e->h_f0.data(), sizeof(double) * N * B, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_L, e->h_L.data(), sizeof(double) * B, cudaMemcpyHostToDevice, e->stream));
  if (!e->uniform) {
    if (!e->d_x) {
      CK(cudaMalloc(&e->d_x, sizeof(double) * N * B));
      CK(cudaMalloc(&e->d_eta_bnd, sizeof(double) * 2 * B));
      CK(cudaMalloc(&e->d_scratch, sizeof(double) * 2 * e->ni * B));
    }
    CK(cudaMemcpyAsync(e->d_x, e->h_x.data(), sizeof(double) * N * B, cudaMemcpyHostToDevice, e->stream));
  }
  e->params_dirty = false;
  return SCFTB_OK;
}
}  // namespace scftb

extern "C" {

const char *scftb_last_error(void) { return g_last_error.c_str(); }

long scftb_launch_count(int reset) {
  long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int scftb_create(const scftb_config *cfg, scftb_engine **out) {
  if (!cfg || !out) return fail(SCFTB_ERR_ARG, "null argument");
  if (cfg->N < 4 || cfg->nsteps < 2 || cfg->max_batch < 1) return fail(SCFTB_ERR_ARG, "N>=4, nsteps>=2, max_batch>=1");
  if (cfg->scheme != SCFTB_IE_ROWSCALE && cfg->scheme != SCFTB_IE_CONSISTENT && cfg->scheme != SCFTB_IRK4_CONSISTENT)
    return fail(SCFTB_ERR_ARG, "unknown scheme");
  scftb_engine *e = new scftb_engine();
  e->cfg = *cfg;
  e->ni = cfg->N - 2;
  e->uniform = true;
  e->params_dirty = true;
  e->last_nprob = 0;
  e->timing = false;
  e->d_x = e->d_eta_bnd = e->d_scratch = nullptr;
  e->d_eta = e->d_out = e->d_phi = e->d_Q = e->d_f0 = e->d_L = e->d_w = e->d_hist = e->d_eta_full = nullptr;
  e->stream = nullptr;
  const int N = cfg->N, B = cfg->max_batch, n = cfg->nsteps;
  if (cfg->quadrature == SCFTB_QUAD_ROMBERG) {
    if (romberg_weights(n, 1.0 / n, e->h_w)) {
      delete e;
      return fail(SCFTB_ERR_ARG, "Romberg quadrature needs nsteps = 2^k >= 16 (romint.c:28-33)");
    }
  } else
    trapezoid_weights(n, 1.0 / n, e->h_w);
  // pair weights for the fused half-history quadrature: the weights are symmetric (w_j = w_{n-j}), so
  // sum_j w_j q_j q_{n-j} = sum_{j>n/2} 2 w_j q_j q_{n-j} + [n even] w_{n/2} q_{n/2}^2
  std::vector<double> wq(e->h_w);
  for (int j = 0; j <= n; j++) wq[j] = (2 * j > n) ? 2.0 * e->h_w[j] : ((2 * j == n) ? e->h_w[j] : 0.0);
  if (choose_any(cfg->scheme, e->ni, true, e->kc, cfg->nsteps, cfg->max_batch)) {
    delete e;
    return fail(SCFTB_ERR_ARG, cfg->scheme == SCFTB_IRK4_CONSISTENT
                                   ? "N too large for the register-resident IRK4 march (N <= 2050 in this build)"
                                   : "N too large for the register-resident march (N <= 4098 in this build)");
  }
#define CKD(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      int rc = fail(SCFTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));           \
      scftb_destroy(e); /* frees whatever was allocated so far (members start out null) */         \
      return rc;                                                                                   \
    }                                                                                              \
  } while (0)
  CKD(cudaSetDevice(cfg->device));
  CKD(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->SL = (size_t)e->kc.T * e->kc.C;
  int sms = 0, occ = 0;
  CKD(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
  if (e->kc.dyn_smem) {
    CKD(cudaFuncSetAttribute((const void *)e->kc.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, e->kc.dyn_smem));
    // four CTAs of ~50 KB each need most of the SM's unified L1/shared array configured as shared memory
    CKD(cudaFuncSetAttribute((const void *)e->kc.fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared));
  }
  CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)e->kc.fn, e->kc.T, e->kc.dyn_smem));
  if (occ < 1) occ = 1;
  // The occupancy query answers 1 for any kernel that allocates tensor memory (CUDA 12.9), but the hardware does keep
  // four such CTAs resident when registers, shared memory and TMEM columns allow it (measured: 592 CTAs of this kernel
  // run as ONE wave, tools/lat2.py).  The slot count only has to be an upper bound on resident CTAs (history buffers are
  // indexed by blockIdx.x), so the designed value is used.
  if (e->kc.max_occ) occ = e->kc.max_occ;
  if (getenv("SCFTB_FORCE_OCC")) occ = atoi(getenv("SCFTB_FORCE_OCC"));   // experiments: resident-CTA slots per SM
  if (getenv("SCFTB_DEBUG")) fprintf(stderr, "scftb: march kernel C=%d T=%d dyn_smem=%d occupancy=%d CTAs/SM\n", e->kc.C, e->kc.T, e->kc.dyn_smem, occ);
  e->slots = std::min(B, sms * occ);
  e->nslices = cfg->store_history ? n + 1 : n / 2 + 1;
  size_t nh = (size_t)(cfg->store_history ? B : e->slots) * e->nslices * e->SL;
  e->h_tau.assign(B, 0.0);
  e->h_L.assign(B, 1.0);
  e->h_f0.assign((size_t)N * B, 1.0);
  CKD(cudaMalloc(&e->d_eta, sizeof(double) * e->ni * B));
  CKD(cudaMalloc(&e->d_out, sizeof(double) * e->ni * B));
  CKD(cudaMalloc(&e->d_phi, sizeof(double) * N * B));
  CKD(cudaMalloc(&e->d_eta_full, sizeof(double) * N * B));
  CKD(cudaMalloc(&e->d_Q, sizeof(double) * B));
  CKD(cudaMalloc(&e->d_f0, sizeof(double) * N * B));
  CKD(cudaMalloc(&e->d_L, sizeof(double) * B));
  CKD(cudaMalloc(&e->d_w, sizeof(double) * (n + 1)));
  CKD(cudaMalloc(&e->d_hist, sizeof(double) * nh));
  CKD(cudaMemcpy(e->d_w, wq.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
  *out = e;
  return SCFTB_OK;
}

int scftb_set_timing(scftb_engine *e, int on) {
  if (!e) return fail(SCFTB_ERR_ARG, "null engine");
  e->timing = on != 0;
  return SCFTB_OK;
}

int scftb_get_march_ms(scftb_engine *e, double *total_ms, int *count) {
  if (!e || !total_ms || !count) return fail(SCFTB_ERR_ARG, "null argument");
  double tot = 0.0;
  int n = 0;
  for (auto &ev : e->ev_pending) {
    CK(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms; n++;
    e->ev_free.push_back(ev);
  }
  e->ev_pending.clear();
  *total_ms = tot; *count = n;
  return SCFTB_OK;
}

int scftb_engine_max_batch(scftb_engine *e) { return e ? e->cfg.max_batch : 0; }

int scftb_get_kernel_name(scftb_engine *e, char *buf, int len) {
  if (!e || !buf || len < 1) return fail(SCFTB_ERR_ARG, "null argument");
  const bool tm = e->kc.fn == (march_fn)march_tm_kernel || e->kc.fn == (march_fn)march_irk4_tm_kernel;
  snprintf(buf, len, "%s<%d,%d,%s>%s", e->cfg.scheme == SCFTB_IRK4_CONSISTENT ? (tm ? "march_irk4_tm_kernel" : "march_irk4_kernel") : (tm ? "march_tm_kernel" : "march_ie_kernel"),
           e->kc.C, e->kc.T, e->uniform ? "uniform" : "mesh", tm ? " (coefficients in tensor memory)" : "");
  return SCFTB_OK;
}

int scftb_get_slots(scftb_engine *e, int *slots) {
  if (!e || !slots) return fail(SCFTB_ERR_ARG, "null argument");
  *slots = e->slots;
  return SCFTB_OK;
}

int scftb_destroy(scftb_engine *e) {
  if (!e) return SCFTB_OK;
  scftb_unbind_engine(e);
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->solver_state && e->solver_state_free) e->solver_state_free(e->solver_state);
  if (e->diblock_state && e->diblock_state_free) e->diblock_state_free(e->diblock_state);
  for (double *p : {e->d_eta, e->d_out, e->d_phi, e->d_Q, e->d_f0, e->d_L, e->d_x, e->d_eta_bnd, e->d_w, e->d_hist,
                    e->d_eta_full, e->d_scratch})
    if (p) cudaFree(p);
  for (cudaEvent_t ev : e->ev_pipe) cudaEventDestroy(ev);
  if (e->ev_last) cudaEventDestroy(e->ev_last);
  for (auto &ev : e->ev_pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto &ev : e->ev_free) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return SCFTB_OK;
}

int scftb_set_problem(scftb_engine *e, int p, double tau, double L, const double *x) {
  if (!e || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  const int N = e->cfg.N, B = e->cfg.max_batch;
  if (x && e->uniform) {  // switch the engine to explicit node coordinates
    e->h_x.assign((size_t)N * B, 0.0);
    for (int q = 0; q < B; q++)
      for (int i = 0; i < N; i++) e->h_x[(size_t)q * N + i] = e->h_L[q] * i / (N - 1);
    e->uniform = false;
    if (choose_any(e->cfg.scheme, e->ni, false, e->kc, e->cfg.nsteps, e->cfg.max_batch)) return fail(SCFTB_ERR_ARG, "N too large");
  }
  std::vector<double> xs(N);
  for (int q = (p < 0 ? 0 : p); q < (p < 0 ? B : p + 1); q++) {
    e->h_tau[q] = tau;
    e->h_L[q] = L;
    for (int i = 0; i < N; i++) xs[i] = x ? x[i] : L * i / (N - 1);  // subdivided_hyper_rectangle, drivescft.cc:94-98
    if (!e->uniform) std::copy(xs.begin(), xs.end(), e->h_x.begin() + (size_t)q * N);
    f0_given(N, xs.data(), tau, &e->h_f0[(size_t)q * N]);
  }
  e->params_dirty = true;
  return SCFTB_OK;
}

}  // extern "C"

namespace scftb {
int order_before_launch(scftb_engine *e, cudaStream_t st) {
  if (e->have_last && e->st_last != st) CK(cudaStreamWaitEvent(st, e->ev_last, 0));
  return SCFTB_OK;
}
int note_launch(scftb_engine *e, cudaStream_t st) {
  if (!e->ev_last) CK(cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming));
  CK(cudaEventRecord(e->ev_last, st));
  e->st_last = st; e->have_last = true;
  return SCFTB_OK;
}

int launch_march(scftb_engine *e, int nprob, const double *d_eta, long long eta_stride, double *d_out,
                 long long out_stride, const int *d_skip, cudaStream_t st, int p0, bool pshare) {
  MarchParams P{};
  const size_t N = e->cfg.N, o = (size_t)p0;
  P.N = e->cfg.N; P.ni = e->ni; P.nsteps = e->cfg.nsteps;
  P.scheme = e->cfg.scheme; P.nprob = nprob; P.store_full = e->cfg.store_history;
  P.uniform = e->uniform ? 1 : 0; P.sign = e->cfg.sign; P.pshare = pshare ? 1 : 0;
  P.eta_mid = d_eta; P.eta_stride = eta_stride; P.out_stride = out_stride; P.skip = d_skip;
  P.f0 = e->d_f0 + o * N; P.L = e->d_L + o; P.w = e->d_w;
  P.x = e->d_x ? e->d_x + o * N : nullptr; P.eta_bnd = e->d_eta_bnd ? e->d_eta_bnd + 2 * o : nullptr;
  P.hist_stride = (long long)e->nslices * (long long)e->SL;
  P.hist = e->d_hist + (e->cfg.store_history ? o * (size_t)P.hist_stride : 0);
  P.out = d_out; P.phi = e->d_phi + o * N; P.Q = e->d_Q + o; P.eta_full = e->d_eta_full + o * N;
  if (!e->uniform) {
    spline_bnd_kernel<<<(nprob + 63) / 64, 64, 0, st>>>(nprob, P.N, P.x, d_eta, eta_stride, e->d_scratch + o * 2 * e->ni,
                                                          e->d_eta_bnd + 2 * o, P.pshare);
    g_launches++;
  }
  int grid = std::min(nprob, e->slots);
  {
    int rc = order_before_launch(e, st);
    if (rc) return rc;
  }
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (e->timing) {
    if (!e->ev_free.empty()) { ev = e->ev_free.back(); e->ev_free.pop_back(); }
    else { CK(cudaEventCreate(&ev.first)); CK(cudaEventCreate(&ev.second)); }
    CK(cudaEventRecord(ev.first, st));
  }
  e->kc.fn<<<grid, e->kc.T, e->kc.dyn_smem, st>>>(P);
  if (e->timing) { CK(cudaEventRecord(ev.second, st)); e->ev_pending.push_back(ev); }
  g_launches++;
  CK(cudaGetLastError());
  e->last_nprob = nprob;
  return note_launch(e, st);
}
}  // namespace scftb

extern "C" {

int scftb_residual_batch_device(scftb_engine *e, int nprob, const double *d_eta_mid, double *d_out, void *stream) {
  if (!e || nprob < 1 || nprob > e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "nprob out of range");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (e->params_dirty) {
    int rc = upload_params(e);
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->stream));
  }
  return launch_march(e, nprob, d_eta_mid, e->ni, d_out, e->ni, nullptr, st);
}

int scftb_residual_batch(scftb_engine *e, int nprob, const double *eta_mid, double *out) {
  if (!e || nprob < 1 || nprob > e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "nprob out of range");
  CK(cudaSetDevice(e->cfg.device));
  int rc = upload_params(e);
  if (rc) return rc;
  const size_t ni = e->ni;
  const int waves = (nprob + e->slots - 1) / e->slots;
  if (waves < 4) {   // small batch: one copy in, one launch, one copy out
    const size_t bytes = sizeof(double) * ni * nprob;
    CK(cudaMemcpyAsync(e->d_eta, eta_mid, bytes, cudaMemcpyHostToDevice, e->stream));
    rc = launch_march(e, nprob, e->d_eta, e->ni, e->d_out, e->ni, nullptr, e->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCFTB_OK;
  }
  // large batch: chunks of whole waves of resident CTAs; the copies of chunk c+1 / c-1 run on their own streams
  // under the march of chunk c (pinned host buffers make them asynchronous; pageable ones still work)
  const int chunk = e->slots * std::max(1, waves / 3);
  const int nchunks = (nprob + chunk - 1) / chunk;
  if (!e->s_in) {
    CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  }
  while ((int)e->ev_pipe.size() < 2 * nchunks) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev_pipe.push_back(ev);
  }
  CK(cudaStreamSynchronize(e->stream));   // parameters uploaded, earlier work on this engine finished
  for (int c = 0; c < nchunks; c++) {
    const int p0 = c * chunk, nb = std::min(chunk, nprob - p0);
    const size_t off = ni * p0, bytes = sizeof(double) * ni * nb;
    CK(cudaMemcpyAsync(e->d_eta + off, eta_mid + off, bytes, cudaMemcpyHostToDevice, e->s_in));
    CK(cudaEventRecord(e->ev_pipe[2 * c], e->s_in));
    CK(cudaStreamWaitEvent(e->stream, e->ev_pipe[2 * c], 0));
    rc = launch_march(e, nb, e->d_eta + off, e->ni, e->d_out + off, e->ni, nullptr, e->stream, p0);
    if (rc) return rc;
    CK(cudaEventRecord(e->ev_pipe[2 * c + 1], e->stream));
    CK(cudaStreamWaitEvent(e->s_out, e->ev_pipe[2 * c + 1], 0));
    CK(cudaMemcpyAsync(out + off, e->d_out + off, bytes, cudaMemcpyDeviceToHost, e->s_out));
  }
  CK(cudaStreamSynchronize(e->s_out));
  CK(cudaStreamSynchronize(e->stream));
  e->last_nprob = nprob;
  return SCFTB_OK;
}

int scftb_residual(scftb_engine *e, const double *eta_mid, double *out) { return scftb_residual_batch(e, 1, eta_mid, out); }

}  // extern "C"

namespace scftb {
int residual_batch_shared(scftb_engine *e, int nprob, const double *eta_mid, double *out) {
  if (!e || nprob < 1 || nprob > e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "nprob out of range");
  CK(cudaSetDevice(e->cfg.device));
  int rc = upload_params(e);
  if (rc) return rc;
  const size_t bytes = sizeof(double) * (size_t)e->ni * nprob;
  CK(cudaMemcpyAsync(e->d_eta, eta_mid, bytes, cudaMemcpyHostToDevice, e->stream));
  rc = launch_march(e, nprob, e->d_eta, e->ni, e->d_out, e->ni, nullptr, e->stream, 0, true);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return SCFTB_OK;
}
}  // namespace scftb

extern "C" {

static int fetch(scftb_engine *e, const double *d, size_t off, size_t cnt, double *h) {
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(h, d + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

int scftb_get_phi(scftb_engine *e, int p, double *phi) {
  if (!e || p < 0 || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  return fetch(e, e->d_phi, (size_t)p * e->cfg.N, e->cfg.N, phi);
}
int scftb_get_Q(scftb_engine *e, int p, double *Q) {
  if (!e || p < 0 || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  return fetch(e, e->d_Q, p, 1, Q);
}
int scftb_get_eta_full(scftb_engine *e, int p, double *eta) {
  if (!e || p < 0 || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  return fetch(e, e->d_eta_full, (size_t)p * e->cfg.N, e->cfg.N, eta);
}
int scftb_get_f0_given(scftb_engine *e, int p, double *f0) {
  if (!e || p < 0 || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  std::copy(e->h_f0.begin() + (size_t)p * e->cfg.N, e->h_f0.begin() + (size_t)(p + 1) * e->cfg.N, f0);
  return SCFTB_OK;
}

int scftb_get_q_history(scftb_engine *e, int p, double *hist) {
  if (!e || p < 0 || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  if (!e->cfg.store_history) return fail(SCFTB_ERR_STATE, "engine created with store_history = 0");
  const int N = e->cfg.N, n = e->cfg.nsteps, T = e->kc.T, C = e->kc.C;
  std::vector<double> raw((size_t)e->nslices * e->SL);
  int rc = fetch(e, e->d_hist, (size_t)p * e->nslices * e->SL, raw.size(), raw.data());
  if (rc) return rc;
  std::fill(hist, hist + (size_t)N * (n + 1), 0.0);
  for (int j = 0; j <= n; j++)
    for (int g = 0; g < e->ni; g++) {
      int t = g / C, k = g % C;
      hist[(size_t)(g + 1) * (n + 1) + j] = raw[(size_t)j * e->SL + hist_index(C, T, t, k)];
    }
  return SCFTB_OK;
}

int scftb_free_energy(scftb_engine *e, int p, double f0bar, double *F) {
  if (!e || p < 0 || p >= e->cfg.max_batch || !F) return fail(SCFTB_ERR_ARG, "bad argument");
  const int N = e->cfg.N;
  std::vector<double> eta(N), xs(N);
  int rc = scftb_get_eta_full(e, p, eta.data());
  if (rc) return rc;
  const double L = e->h_L[p], tau = e->h_tau[p];
  for (int i = 0; i < N; i++) xs[i] = e->uniform ? L * i / (N - 1) : e->h_x[(size_t)p * N + i];
  std::vector<double> w;
  if (f0bar <= 0) {  // testFiBar.cc:19-50
    const int M = (1 << 16) + 1;
    std::vector<double> xx(M), ff(M);
    for (int i = 0; i < M; i++) xx[i] = L / (M - 1) * i;
    f0_given(M, xx.data(), tau, ff.data());
    romberg_weights(M - 1, L / (M - 1), w);
    double s = 0;
    for (int i = 0; i < M; i++) s += w[i] * ff[i];
    f0bar = s / L;
  }
  const int nplot = (1 << 18) + 1;  // scft.cc:271
  std::vector<double> xp(nplot), f0(nplot);
  for (int i = 0; i < nplot; i++) xp[i] = L * i / (nplot - 1);
  f0_given(nplot, xp.data(), tau, f0.data());
  romberg_weights(nplot - 1, L / (nplot - 1), w);
  double I = 0;
  int k = 0;
  for (int i = 0; i < nplot; i++) {  // piecewise-linear eta_h (FEFieldFunction on Q1, scft.cc:280-281)
    while (k < N - 2 && xp[i] > xs[k + 1]) k++;
    double t = (xp[i] - xs[k]) / (xs[k + 1] - xs[k]);
    I += w[i] * ((1 - t) * eta[k] + t * eta[k + 1]) * f0[i];
  }
  *F = (I / f0bar / L + std::log(f0bar)) / (-1000.);  // scft.cc:446-447
  return SCFTB_OK;
}

}  // extern "C"
