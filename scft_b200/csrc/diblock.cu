// diblock.cu — two-species (AB diblock) extension of the residual: q and q+ marched as separate sweeps
// (SURVEY.md section 8(f)-4).  Not in the reference, whose melt is one species with q+(x,s) = q(x,1-s)
// (drivescft.cc:189-190); parity is pinned by oracle/scft_oracle.c::orc_residual_ab only.
// The kernel is the TWO = true instantiation of march_ie_kernel (march1d.cuh): same substructured tridiagonal
// solve, four segments (q through block A, block B; q+ through block B, block A), every q slice kept in HBM
// (8 B written + 8 B read per pair of propagator DOF-steps), phi_A / phi_B accumulated by the q+ sweep.
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.h"
#include "march1d.cuh"

using namespace scftb;

namespace {

struct DiblockState {
  int jf = 0;
  double fA = 0.0;
  KernelChoice kc{};
  int slots = 0;
  size_t SL = 0;
  bool uniform_kernel = true;
  std::vector<double> h_chi;
  bool chi_dirty = true;
  double *d_chi = nullptr, *d_wA = nullptr, *d_wB = nullptr, *d_phiB = nullptr, *d_w = nullptr, *d_out = nullptr,
         *d_hist = nullptr, *d_eta_bndB = nullptr, *d_scratch = nullptr;
  void release() {
    for (double *p : {d_chi, d_wA, d_wB, d_phiB, d_w, d_out, d_hist, d_eta_bndB, d_scratch})
      if (p) cudaFree(p);
  }
};
void free_diblock_state(void *p) {
  DiblockState *s = (DiblockState *)p;
  s->release();
  delete s;
}

template <int C, int T, int MINB>
march_fn pick_ab(bool uni) {
  return uni ? (march_fn)march_ie_kernel<C, T, true, MINB, false, true> : (march_fn)march_ie_kernel<C, T, false, MINB, false, true>;
}
// same shapes as choose_kernel (engine.cu); the two-species instantiation carries a second accumulator per node
int choose_kernel_ab(int ni, bool uni, KernelChoice &kc) {
  int C = 1;
  while (C < 16 && (ni + C - 1) / C > 128) C *= 2;
  const int need = (ni + C - 1) / C;
  if (need > 256) return 1;
  const int T = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 256));
  kc.fn = nullptr;
  if (C == 1 && T == 32) kc.fn = pick_ab<1, 32, 8>(uni);
  if (C == 1 && T == 64) kc.fn = pick_ab<1, 64, 6>(uni);
  if (C == 1 && T == 128) kc.fn = pick_ab<1, 128, 4>(uni);
  if (C == 2 && T == 128) kc.fn = pick_ab<2, 128, 4>(uni);
  if (C == 4 && T == 128) kc.fn = pick_ab<4, 128, 4>(uni);
  if (C == 8 && T == 128) kc.fn = pick_ab<8, 128, 3>(uni);
  if (C == 16 && T == 128) kc.fn = pick_ab<16, 128, 1>(uni);
  if (C == 16 && T == 256) kc.fn = pick_ab<16, 256, 1>(uni);
  if (!kc.fn) return 1;
  kc.C = C; kc.T = T;
  return 0;
}

// block quadrature weights over contour indices lo..lo+m (zero elsewhere): the engine's Romberg rule when the
// block has 2^k >= 16 steps (romint.c:28-33), else the trapezoid rule (simple_FEM_1D_transient.m:120-124)
void block_weights(int quadrature, int n, int lo, int m, std::vector<double> &w) {
  std::vector<double> wb;
  if (quadrature != SCFTB_QUAD_ROMBERG || romberg_weights(m, 1.0 / n, wb)) trapezoid_weights(m, 1.0 / n, wb);
  w.assign(n + 1, 0.0);
  for (int j = 0; j <= m; j++) w[lo + j] = wb[j];
}

int ensure_state(scftb_engine *e, double fA, DiblockState **out) {
  if (e->cfg.scheme == SCFTB_IRK4_CONSISTENT) return fail(SCFTB_ERR_ARG, "two-species march: implicit-Euler schemes only");
  const int n = e->cfg.nsteps, N = e->cfg.N, B = e->cfg.max_batch;
  const double jfd = fA * n;
  const int jf = (int)std::llround(jfd);
  if (!(fA > 0.0 && fA < 1.0) || jf < 1 || jf >= n || std::fabs(jfd - jf) > 1e-9 * n)
    return fail(SCFTB_ERR_ARG, "fA * nsteps must be an integer number of contour steps in (0, nsteps)");
  DiblockState *s = (DiblockState *)e->diblock_state;
  if (!s) {
    s = new DiblockState();
    e->diblock_state = s;
    e->diblock_state_free = free_diblock_state;
    s->h_chi.assign(B, 0.0);
  }
  CK(cudaSetDevice(e->cfg.device));
  if (!s->d_chi) {
    CK(cudaMalloc(&s->d_chi, sizeof(double) * B));
    CK(cudaMalloc(&s->d_wA, sizeof(double) * (n + 1)));
    CK(cudaMalloc(&s->d_wB, sizeof(double) * (n + 1)));
    CK(cudaMalloc(&s->d_phiB, sizeof(double) * N * B));
    CK(cudaMalloc(&s->d_w, sizeof(double) * 2 * e->ni * B));
    CK(cudaMalloc(&s->d_out, sizeof(double) * 2 * e->ni * B));
  }
  if (jf != s->jf) {
    std::vector<double> wA, wB;
    block_weights(e->cfg.quadrature, n, 0, jf, wA);
    block_weights(e->cfg.quadrature, n, jf, n - jf, wB);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(s->d_wA, wA.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_wB, wB.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
    s->jf = jf; s->fA = fA;
  }
  *out = s;
  return SCFTB_OK;
}

// kernel choice and history (all n+1 slices per resident CTA) follow the engine's current mesh mode
int ensure_kernel(scftb_engine *e, DiblockState *s) {
  if (s->kc.fn && s->uniform_kernel == e->uniform) return SCFTB_OK;
  if (choose_kernel_ab(e->ni, e->uniform, s->kc)) return fail(SCFTB_ERR_ARG, "N too large for the two-species march");
  s->uniform_kernel = e->uniform;
  int sms = 0, occ = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->cfg.device));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)s->kc.fn, s->kc.T, 0));
  if (occ < 1) occ = 1;
  const int slots = std::min(e->cfg.max_batch, sms * occ);
  const size_t SL = (size_t)s->kc.T * s->kc.C;
  if (slots != s->slots || SL != s->SL || !s->d_hist) {
    if (s->d_hist) { CK(cudaStreamSynchronize(e->stream)); CK(cudaFree(s->d_hist)); s->d_hist = nullptr; }
    CK(cudaMalloc(&s->d_hist, sizeof(double) * (size_t)slots * (e->cfg.nsteps + 1) * SL));
    s->slots = slots; s->SL = SL;
  }
  return SCFTB_OK;
}

int launch_ab(scftb_engine *e, DiblockState *s, int nprob, const double *d_w, double *d_out, cudaStream_t st, bool pshare = false,
              long long w_stride = 0, long long out_stride = 0, const int *d_skip = nullptr) {
  int rc = ensure_kernel(e, s);
  if (rc) return rc;
  if (s->chi_dirty) {
    CK(cudaMemcpyAsync(s->d_chi, s->h_chi.data(), sizeof(double) * e->cfg.max_batch, cudaMemcpyHostToDevice, st));
    s->chi_dirty = false;
  }
  MarchParams P{};
  P.N = e->cfg.N; P.ni = e->ni; P.nsteps = e->cfg.nsteps;
  P.scheme = e->cfg.scheme; P.nprob = nprob; P.store_full = 0;
  P.uniform = e->uniform ? 1 : 0; P.sign = e->cfg.sign; P.pshare = pshare ? 1 : 0;
  P.eta_mid = d_w; P.eta_stride = w_stride ? w_stride : 2 * (long long)e->ni;
  P.out_stride = out_stride ? out_stride : 2 * (long long)e->ni; P.skip = d_skip;
  P.f0 = e->d_f0; P.L = e->d_L; P.x = e->d_x; P.eta_bnd = e->d_eta_bnd; P.w = e->d_w;
  P.hist = s->d_hist; P.hist_stride = (long long)(e->cfg.nsteps + 1) * (long long)s->SL;
  P.out = d_out; P.phi = e->d_phi; P.Q = e->d_Q; P.eta_full = nullptr;
  P.jf = s->jf; P.wA = s->d_wA; P.wB = s->d_wB; P.chi = s->d_chi; P.phiB = s->d_phiB; P.eta_bndB = s->d_eta_bndB;
  if (!e->uniform) {   // natural-spline wall values of both fields (scft.cc:452-490)
    if (!s->d_eta_bndB) {
      CK(cudaMalloc(&s->d_eta_bndB, sizeof(double) * 2 * e->cfg.max_batch));
      CK(cudaMalloc(&s->d_scratch, sizeof(double) * 2 * e->ni * e->cfg.max_batch));
      P.eta_bndB = s->d_eta_bndB;
    }
    spline_bnd_launch(nprob, P.N, e->d_x, d_w, P.eta_stride, e->d_scratch, e->d_eta_bnd, st, P.pshare);
    spline_bnd_launch(nprob, P.N, e->d_x, d_w + e->ni, P.eta_stride, s->d_scratch, s->d_eta_bndB, st, P.pshare);
  }
  const int grid = std::min(nprob, s->slots);
  if ((rc = order_before_launch(e, st))) return rc;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (e->timing) {   // same per-launch CUDA-event timing as the one-sweep march (scftb_set_timing / scftb_get_march_ms)
    if (!e->ev_free.empty()) { ev = e->ev_free.back(); e->ev_free.pop_back(); }
    else { CK(cudaEventCreate(&ev.first)); CK(cudaEventCreate(&ev.second)); }
    CK(cudaEventRecord(ev.first, st));
  }
  s->kc.fn<<<grid, s->kc.T, 0, st>>>(P);
  if (e->timing) { CK(cudaEventRecord(ev.second, st)); e->ev_pending.push_back(ev); }
  g_launches++;
  CK(cudaGetLastError());
  return note_launch(e, st);
}

}  // namespace

extern "C" {

int scftb_set_diblock(scftb_engine *e, int p, double fA, double chiN) {
  if (!e || p >= e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "bad problem index");
  DiblockState *s = nullptr;
  int rc = ensure_state(e, fA, &s);
  if (rc) return rc;
  for (int q = (p < 0 ? 0 : p); q < (p < 0 ? e->cfg.max_batch : p + 1); q++) s->h_chi[q] = chiN;
  s->chi_dirty = true;
  return SCFTB_OK;
}

static int residual_ab_impl(scftb_engine *e, int nprob, const double *w, double *out, bool pshare) {
  if (!e || !w || !out || nprob < 1 || nprob > e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "nprob out of range");
  DiblockState *s = (DiblockState *)e->diblock_state;
  if (!s) return fail(SCFTB_ERR_STATE, "scftb_set_diblock has not been called on this engine");
  CK(cudaSetDevice(e->cfg.device));
  int rc = upload_params(e);
  if (rc) return rc;
  const size_t bytes = sizeof(double) * 2 * (size_t)e->ni * nprob;
  CK(cudaMemcpyAsync(s->d_w, w, bytes, cudaMemcpyHostToDevice, e->stream));
  rc = launch_ab(e, s, nprob, s->d_w, s->d_out, e->stream, pshare);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, s->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return SCFTB_OK;
}

int scftb_residual_ab_batch(scftb_engine *e, int nprob, const double *w, double *out) {
  return residual_ab_impl(e, nprob, w, out, false);
}
}  // extern "C"
namespace scftb {
int residual_ab_batch_shared(scftb_engine *e, int nprob, const double *w, double *out) {
  return residual_ab_impl(e, nprob, w, out, true);
}
}  // namespace scftb
extern "C" {

int scftb_residual_ab(scftb_engine *e, const double *w, double *out) { return scftb_residual_ab_batch(e, 1, w, out); }
}  // extern "C"
namespace scftb {
// device-pointer launch with caller strides and skip flags, for the device-resident mixers (mixer.cu)
int launch_residual_ab(scftb_engine *e, int nprob, const double *d_w, long long w_stride, double *d_out, long long out_stride,
                       const int *d_skip, cudaStream_t st) {
  DiblockState *s = (DiblockState *)e->diblock_state;
  if (!s) return fail(SCFTB_ERR_STATE, "scftb_set_diblock has not been called on this engine");
  return launch_ab(e, s, nprob, d_w, d_out, st, false, w_stride, out_stride, d_skip);
}
}  // namespace scftb
extern "C" {

int scftb_get_phi_ab(scftb_engine *e, int p, double *phiA, double *phiB) {
  if (!e || p < 0 || p >= e->cfg.max_batch || !phiA || !phiB) return fail(SCFTB_ERR_ARG, "bad argument");
  DiblockState *s = (DiblockState *)e->diblock_state;
  if (!s) return fail(SCFTB_ERR_STATE, "scftb_set_diblock has not been called on this engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->stream));
  const size_t N = e->cfg.N;
  CK(cudaMemcpy(phiA, e->d_phi + (size_t)p * N, sizeof(double) * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(phiB, s->d_phiB + (size_t)p * N, sizeof(double) * N, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

}  // extern "C"
