// engine.h — internal declarations shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <utility>
#include <string>
#include <vector>

#include "../../include/scft_b200.h"

namespace scftb {
int fail(int code, const std::string &msg);
int romberg_weights(int m, double hh, std::vector<double> &w);
void trapezoid_weights(int m, double hh, std::vector<double> &w);
void f0_given(int N, const double *x, double tau, double *f0);
int gauss_jordan(std::vector<double> &a, int n, std::vector<double> &b, bool nudge);
}  // namespace scftb

namespace scftb {
struct MarchParams;
typedef void (*march_fn)(MarchParams);
struct KernelChoice {
  int C, T;
  march_fn fn;
  int dyn_smem = 0;   // dynamic shared memory per CTA (the tensor-memory kernel pads itself to four CTAs per SM)
  int max_occ = 0;    // cap on resident CTAs per SM (0: whatever the occupancy query says)
};
extern std::atomic<long> g_launches;
int choose_kernel(int ni, bool uni, KernelChoice &kc, bool odd);
void spline_bnd_launch(int nprob, int N, const double *x, const double *eta_mid, long long eta_stride, double *scratch,
                       double *eta_bnd, cudaStream_t st, int pshare = 0);
}  // namespace scftb

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return scftb::fail(SCFTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));             \
  } while (0)


struct scftb_engine {
  scftb_config cfg;
  int ni;
  bool uniform;
  scftb::KernelChoice kc;
  int slots;       // resident CTAs
  int nslices;     // history slices per problem/slot
  size_t SL;       // doubles per slice
  cudaStream_t stream;
  // host mirrors
  std::vector<double> h_tau, h_L, h_x, h_f0, h_w;
  bool params_dirty;
  int last_nprob;
  // device buffers
  double *d_eta, *d_out, *d_phi, *d_Q, *d_f0, *d_L, *d_x, *d_eta_bnd, *d_w, *d_hist, *d_eta_full, *d_scratch;
  // lean history lives per resident CTA slot of THIS engine, so two march launches of one engine must never overlap:
  // every launch records ev_last on its stream and a launch on a different stream first waits on it
  cudaEvent_t ev_last = nullptr;
  cudaStream_t st_last = nullptr;
  bool have_last = false;
  // optional per-launch timing of the march kernel (CUDA events on the launching stream)
  bool timing;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
  // copy streams + events of the pipelined host-buffer batch call (scftb_residual_batch), created on first use
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_pipe;
  // state a device-resident solver keeps between calls on this engine (Broyden's QR factors for jc), with its deleter
  void *solver_state = nullptr;
  void (*solver_state_free)(void *) = nullptr;
  // two-species extension (diblock.cu): its buffers, with their deleter
  void *diblock_state = nullptr;
  void (*diblock_state_free)(void *) = nullptr;
};

namespace scftb {
// launch one batch of residual evaluations; eta/out are device pointers with the given problem strides
// p0: index of the first problem of this launch in the engine's per-problem arrays (d_eta / d_out already point at it)
// pshare: every problem of the launch is evaluated with the parameters (tau, L, mesh, phi_0) of slot p0 and only `out` is
// written — the n perturbed fields of a finite-difference Jacobian (fdjac.c:18-34) of ONE problem, whatever the other
// slots of the engine hold
int launch_march(scftb_engine *e, int nprob, const double *d_eta, long long eta_stride, double *d_out,
                 long long out_stride, const int *d_skip, cudaStream_t st, int p0 = 0, bool pshare = false);
// host-buffer batch of nprob fields of problem 0 (pshare launch), for the host-flow solvers
int residual_batch_shared(scftb_engine *e, int nprob, const double *eta_mid, double *out);
int residual_ab_batch_shared(scftb_engine *e, int nprob, const double *w, double *out);
int launch_residual_ab(scftb_engine *e, int nprob, const double *d_w, long long w_stride, double *d_out, long long out_stride,
                       const int *d_skip, cudaStream_t st);
// unknowns per problem: N-2 (one species) or 2(N-2) (after scftb_set_diblock: eta_A, eta_B)
inline int unknowns(const scftb_engine *e) { return e->diblock_state ? 2 * e->ni : e->ni; }
int upload_params(scftb_engine *e);
// order a march launch on `st` after the engine's previous march launch (no-op on the same stream) / note it
int order_before_launch(scftb_engine *e, cudaStream_t st);
int note_launch(scftb_engine *e, cudaStream_t st);
}  // namespace scftb

// internal accessor (not part of the public ABI)
extern "C" int scftb_engine_max_batch(scftb_engine *e);
extern "C" void scftb_unbind_engine(scftb_engine *e);
