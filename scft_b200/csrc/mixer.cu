// mixer.cu — device-resident batched Anderson mixing (adm_chen semantics, ADM_chen_C.c:18-147).
//
// Every problem of a sweep carries its own iterate/residual history ring, Gram matrix U, vector V,
// relaxation lk and restart index on the device.  One iteration is two launches:
//     march kernel      Y_k = F(X_k)                 (reads/writes the rings in place)
//     anderson_kernel   err_k, U, V, gaussj, X_{k+1}  (one CTA per problem)
// and nothing crosses PCIe between iterations except, when the caller asks, the per-problem
// error norms.  The arithmetic follows the reference operation by operation so that iterates can
// be compared with the CPU path at equal iteration count:
//   * inner products U_ij, V_i are accumulated sequentially in t with separate multiply and add
//     (no fma contraction, no tree reduction); parallelism is across the entries (i,j);
//   * gaussj is the full-pivoting Gauss-Jordan of DEALII_SCFT/src/gaussj.c:7-78 with the same pivot
//     order (row-major scan, ties resolved to the LAST maximum); rows are updated in parallel,
//     which does not change any rounding;
//   * the update X_{k+1} = X_k + sum_j V_j (X_{k-j-1}-X_k) + (1-lk)(Y_k + sum_j V_j (Y_{k-j-1}-Y_k))
//     is evaluated per node in the reference's order (ADM_chen_C.c:114-123).
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.h"

namespace scftb {

constexpr int AND_THREADS = 256;
constexpr int AND_TILE = 64;
constexpr int AND_EPT = 6;    // entries of (U,V) per thread: nn <= 50 -> 50*51/2+50 = 1325 <= 6*256
constexpr int AND_NN_MAX = 50;

struct AndersonParams {
  int n, nprob, R, nm, k, Final, freeze;
  int adm;             // 1: adm.c semantics (see scftb_adm_mixer_create): d = (x + F) - x, window min(k, nm), nudged gaussj, relax = lambda_k
  double relax;        // adm: lambda of this iteration (adm.c:140,151), computed on the host with the host's pow()
  double tol, lmd;
  double *X, *Y;       // [nprob][R][n]
  double *xfinal;      // [nprob][n]
  double *lk, *err;    // [nprob]
  int *k_restart, *done, *iters;
};

__global__ void __launch_bounds__(AND_THREADS) anderson_kernel(AndersonParams A) {
  const int p = blockIdx.x, tid = threadIdx.x, n = A.n, R = A.R, k = A.k;
  if (A.freeze && A.done[p]) return;
  extern __shared__ double sm[];
  const int nm = A.nm;
  double *U = sm;                          // [nm][nm]
  double *V = U + nm * nm;                 // [nm]
  double *dum = V + nm;                    // [nm]
  double *D = dum + nm;                    // [nm+1][AND_TILE+1]  (row nm = Y_k)
  double *red = D + (nm + 1) * (AND_TILE + 1);   // [AND_THREADS]
  int *ired = (int *)(red + AND_THREADS);         // [AND_THREADS]
  int *ipiv = ired + AND_THREADS;                 // [nm]
  __shared__ int s_irow, s_icol, s_sing;
  __shared__ double s_pivinv;

  const double *Xp = A.X + (size_t)p * R * n, *Yp = A.Y + (size_t)p * R * n;
  const double *Yk = Yp + (size_t)(k % R) * n, *Xk = Xp + (size_t)(k % R) * n;

  if (A.adm) {   // adm works on x = f(x) with f(x) = x + F(x) (drivescft.cc:212 comment): d = xnew - x = (x + F) - x, adm.c:163
    double *Yw = A.Y + (size_t)p * R * n + (size_t)(k % R) * n;
    for (int i = tid; i < n; i += AND_THREADS) Yw[i] = __dsub_rn(__dadd_rn(Yw[i], Xk[i]), Xk[i]);
    __syncthreads();
  }

  // ---- err = max |Y_k|, NaN check (ADM_chen_C.c:58-69)
  double e = 0.0;
  int bad = 0;
  for (int i = tid; i < n; i += AND_THREADS) {
    double y = Yk[i];
    if (isnan(y)) bad = 1;
    else if (fabs(y) >= e) e = fabs(y);
  }
  red[tid] = e; ired[tid] = bad;
  __syncthreads();
  for (int s = AND_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) { red[tid] = fmax(red[tid], red[tid + s]); ired[tid] |= ired[tid + s]; }
    __syncthreads();
  }
  const double err = red[0];
  bad = ired[0];
  __syncthreads();
  if (tid == 0) A.err[p] = bad ? nan("") : err;
  if (bad) {   // the reference exit(1)s here; frozen unless the caller asked to keep evaluating (benchmarks)
    if (A.done[p] == 0) {   // keep the field whose residual first went NaN: scftb_mixer_get_x returns it
      for (int i = tid; i < n; i += AND_THREADS) A.xfinal[(size_t)p * n + i] = Xk[i];
      __syncthreads();
      if (tid == 0) { A.done[p] = 2; A.iters[p] = k; }
    }
    if (A.freeze) return;
  }
  if (err < A.tol) {                       // converged: x_old = X[k] (ADM_chen_C.c:71-84)
    for (int i = tid; i < n; i += AND_THREADS) A.xfinal[(size_t)p * n + i] = Xk[i];
    if (tid == 0) { A.done[p] = 1; A.iters[p] = k; }
    return;
  }
  double lk = A.lk[p];
  int m = A.adm ? min(nm, k) : min(nm, k - A.k_restart[p]);   // adm.c:150 nr = IMIN(its-1, NRMAX), its = k+1

  if (m > 0) {
    // ---- U (upper triangle incl. diagonal) and V, sequential in t per entry (ADM_chen_C.c:89-101)
    const int E1 = m * (m + 1) / 2, E = E1 + m;
    double acc[AND_EPT];
    int ei[AND_EPT], ej[AND_EPT];
#pragma unroll
    for (int q = 0; q < AND_EPT; q++) {
      acc[q] = 0.0;
      int eidx = tid + q * AND_THREADS;
      ei[q] = -1; ej[q] = 0;
      if (eidx < E1) {      // invert the triangular numbering: row i holds m-i entries
        int i = 0, rem = eidx;
        while (rem >= m - i) { rem -= m - i; i++; }
        ei[q] = i; ej[q] = i + rem;
      } else if (eidx < E) { ei[q] = eidx - E1; ej[q] = m; }   // V_i = <D_i, Y_k>: row m of D is Y_k
    }
    for (int t0 = 0; t0 < n; t0 += AND_TILE) {
      const int len = min(AND_TILE, n - t0);
      for (int idx = tid; idx < (m + 1) * len; idx += AND_THREADS) {
        int i = idx / len, tt = idx - i * len;
        double yk = Yk[t0 + tt];
        D[i * (AND_TILE + 1) + tt] = (i < m) ? __dsub_rn(yk, Yp[(size_t)((k - i - 1) % R) * n + t0 + tt]) : yk;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < AND_EPT; q++)
        if (ei[q] >= 0) {
          const double *di = D + ei[q] * (AND_TILE + 1), *dj = D + ej[q] * (AND_TILE + 1);
          double a = acc[q];
          for (int tt = 0; tt < len; tt++) a = __dadd_rn(a, __dmul_rn(di[tt], dj[tt]));
          acc[q] = a;
        }
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < AND_EPT; q++)
      if (ei[q] >= 0) {
        if (ej[q] == m) V[ei[q]] = acc[q];
        else { U[ei[q] * m + ej[q]] = acc[q]; U[ej[q] * m + ei[q]] = acc[q]; }
      }
    if (tid < m) ipiv[tid] = 0;
    if (tid == 0) s_sing = 0;
    __syncthreads();

    // ---- gaussj, full pivoting (gaussj.c:19-73); only the solution V is needed afterwards
    for (int it = 0; it < m; it++) {
      double big = -1.0;
      int bidx = -1;
      for (int idx = tid; idx < m * m; idx += AND_THREADS) {
        int j = idx / m, kk = idx - j * m;
        if (ipiv[j] != 1 && ipiv[kk] == 0) {
          double v = fabs(U[idx]);
          if (v >= big) { big = v; bidx = idx; }   // ascending idx per thread: '>=' keeps the last
        }
      }
      red[tid] = big; ired[tid] = bidx;
      __syncthreads();
      for (int s = AND_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
          double vo = red[tid + s]; int io = ired[tid + s];
          if (io >= 0 && (vo > red[tid] || (vo == red[tid] && io > ired[tid]))) { red[tid] = vo; ired[tid] = io; }
        }
        __syncthreads();
      }
      if (tid == 0) {
        // NaN entries never satisfy >=; if nothing was selected the reference keeps its previous
        // irow/icol — treat as singular instead of reading garbage
        int idx = ired[0];
        if (idx < 0) { s_sing = 1; idx = 0; }
        s_irow = idx / m; s_icol = idx - (idx / m) * m;
        ipiv[s_icol]++;
      }
      __syncthreads();
      const int irow = s_irow, icol = s_icol;
      if (s_sing) break;
      if (irow != icol) {
        for (int l = tid; l < m; l += AND_THREADS) { double tmp = U[irow * m + l]; U[irow * m + l] = U[icol * m + l]; U[icol * m + l] = tmp; }
        if (tid == 0) { double tmp = V[irow]; V[irow] = V[icol]; V[icol] = tmp; }
        __syncthreads();
      }
      if (tid == 0) {
        double piv = U[icol * m + icol];
        if (piv == 0.0 && A.adm) piv += 1e-18;    // root gaussj.c:38 (the one adm.c:278 links)
        if (piv == 0.0) s_sing = 1;               // DEALII_SCFT/src/gaussj.c:46-50
        else { s_pivinv = 1.0 / piv; U[icol * m + icol] = 1.0; }
      }
      __syncthreads();
      if (s_sing) break;
      const double pivinv = s_pivinv;
      for (int l = tid; l < m; l += AND_THREADS) U[icol * m + l] = __dmul_rn(U[icol * m + l], pivinv);
      if (tid == 0) V[icol] = __dmul_rn(V[icol], pivinv);
      for (int ll = tid; ll < m; ll += AND_THREADS) dum[ll] = U[ll * m + icol];
      __syncthreads();
      for (int ll = tid; ll < m; ll += AND_THREADS) if (ll != icol) U[ll * m + icol] = 0.0;
      __syncthreads();
      for (int idx = tid; idx < m * m; idx += AND_THREADS) {
        int ll = idx / m, l = idx - ll * m;
        if (ll != icol) U[idx] = __dsub_rn(U[idx], __dmul_rn(U[icol * m + l], dum[ll]));
      }
      for (int ll = tid; ll < m; ll += AND_THREADS) if (ll != icol) V[ll] = __dsub_rn(V[ll], __dmul_rn(V[icol], dum[ll]));
      __syncthreads();
    }
    if (s_sing) {                         // restart with an empty history (ADM_chen_C.c:106-112)
      m = 0;
      lk = A.lmd;
      if (tid == 0) A.k_restart[p] = k;
    }
  }
  __syncthreads();

  // ---- X_{k+1} (ADM_chen_C.c:114-123)
  double *Xn = A.X + (size_t)p * R * n + (size_t)((k + 1) % R) * n;
  const double oml = A.adm ? A.relax : 1 - lk;
  for (int i = tid; i < n; i += AND_THREADS) {
    const double xk = Xk[i], yk = Yk[i];
    double cx = 0.0, cd = 0.0;
    for (int j = 0; j < m; j++) {
      const size_t s = (size_t)((k - j - 1) % R) * n + i;
      cx = __dadd_rn(cx, __dmul_rn(V[j], __dsub_rn(Xp[s], xk)));
      cd = __dadd_rn(cd, __dmul_rn(V[j], __dsub_rn(Yp[s], yk)));
    }
    Xn[i] = __dadd_rn(__dadd_rn(xk, cx), __dmul_rn(oml, __dadd_rn(yk, cd)));
  }
  if (tid == 0 && !A.adm) {               // ADM_chen_C.c:125-132
    if (err < 0.03 && k > 100) lk *= A.lmd;
    if (!A.Final && lk < 1e-5) lk = A.lmd;
    if (A.Final && lk < 1e-15) lk = A.lmd;
    A.lk[p] = lk;
  }
}

}  // namespace scftb

using namespace scftb;

struct scftb_mixer {
  scftb_engine *e;
  int adm = 0;
  int nprob, nn, nm, R, Final, k, freeze, n;
  double lmd, tol;
  double *X, *Y, *xfinal, *lk, *err;
  int *k_restart, *done, *iters;
  size_t smem;
};

extern "C" {

int scftb_mixer_destroy(scftb_mixer *m) {
  if (!m) return SCFTB_OK;
  cudaSetDevice(m->e->cfg.device);
  for (void *p : {(void *)m->X, (void *)m->Y, (void *)m->xfinal, (void *)m->lk, (void *)m->err, (void *)m->k_restart,
                  (void *)m->done, (void *)m->iters})
    if (p) cudaFree(p);
  delete m;
  return SCFTB_OK;
}

int scftb_mixer_create(scftb_engine *e, int nprob, double tol, double lmd, int nn, int Final, scftb_mixer **out) {
  if (!e || !out || nprob < 1 || nprob > e->cfg.max_batch) return fail(SCFTB_ERR_ARG, "mixer: bad argument");
  if (nn < 0 || nn > AND_NN_MAX) return fail(SCFTB_ERR_ARG, "mixer: mixing window nn must be in [0, 50]");
  scftb_mixer *m = new scftb_mixer();
  m->e = e; m->nprob = nprob; m->nn = nn; m->nm = std::min(nn, unknowns(e)); m->R = m->nm + 2;
  m->n = unknowns(e);   // a two-species engine (scftb_set_diblock called before) mixes (eta_A, eta_B) as one vector
  m->Final = Final; m->k = 0; m->lmd = lmd; m->tol = tol; m->freeze = 1;
  m->X = m->Y = m->xfinal = m->lk = m->err = nullptr; m->k_restart = m->done = m->iters = nullptr;
  const size_t n = m->n, ring = (size_t)nprob * m->R * n;
  CK(cudaSetDevice(e->cfg.device));
#define CKM(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { int rc = fail(SCFTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); scftb_mixer_destroy(m); return rc; } } while (0)
  CKM(cudaMalloc(&m->X, sizeof(double) * ring));
  CKM(cudaMalloc(&m->Y, sizeof(double) * ring));
  CKM(cudaMalloc(&m->xfinal, sizeof(double) * nprob * n));
  CKM(cudaMalloc(&m->lk, sizeof(double) * nprob));
  CKM(cudaMalloc(&m->err, sizeof(double) * nprob));
  CKM(cudaMalloc(&m->k_restart, sizeof(int) * nprob));
  CKM(cudaMalloc(&m->done, sizeof(int) * nprob));
  CKM(cudaMalloc(&m->iters, sizeof(int) * nprob));
  auto smem_for = [](size_t nm) {
    return sizeof(double) * (nm * nm + 2 * nm + (nm + 1) * (AND_TILE + 1) + AND_THREADS) + sizeof(int) * (AND_THREADS + nm + 4);
  };
  m->smem = smem_for((size_t)m->nm);
  // the limit is a property of the kernel, shared by every mixer of the process (other windows, other host threads):
  // always set it to the requirement of the largest window, never to this mixer's own
  CKM(cudaFuncSetAttribute((const void *)anderson_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_for((size_t)AND_NN_MAX)));
  *out = m;
  return SCFTB_OK;
}

int scftb_mixer_set_freeze(scftb_mixer *m, int freeze) {
  if (!m) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  m->freeze = freeze != 0;
  return SCFTB_OK;
}

// adm (adm.c:24-313) as a device-resident batched mixer: history ring of NRMAX = 10 (adm.c:6), TOLF = 1e-10 (adm.c:29),
// lambda = 0.05 then 1 - 0.95^its, gaussj with the zero-pivot nudge of the root gaussj.c:38.  The fixed-point map is
// x -> x + (phi0 - phi) (scftb_callback_fixedpoint_c0).  Same iterate/status/get_x calls as the adm_chen mixer.
int scftb_adm_mixer_create(scftb_engine *e, int nprob, scftb_mixer **out) {
  int rc = scftb_mixer_create(e, nprob, 1e-10, 0.0, 10, 0, out);
  if (rc) return rc;
  (*out)->adm = 1;
  return SCFTB_OK;
}

// (re)start from fields x[nprob][n]; device = 1: x is a device pointer, copied on `stream`
int scftb_mixer_reset(scftb_mixer *m, const double *x, int device, void *stream) {
  if (!m || !x) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  scftb_engine *e = m->e;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = device ? (cudaStream_t)stream : e->stream;
  const size_t n = m->n;
  if ((int)n != unknowns(e)) return fail(SCFTB_ERR_STATE, "mixer: the engine changed between one and two species after the mixer was created");
  CK(cudaMemcpy2DAsync(m->X, sizeof(double) * m->R * n, x, sizeof(double) * n, sizeof(double) * n, m->nprob,
                       device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  std::vector<double> lk(m->nprob, m->lmd);
  CK(cudaMemcpyAsync(m->lk, lk.data(), sizeof(double) * m->nprob, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(m->k_restart, 0, sizeof(int) * m->nprob, st));
  CK(cudaMemsetAsync(m->done, 0, sizeof(int) * m->nprob, st));
  CK(cudaMemsetAsync(m->iters, 0, sizeof(int) * m->nprob, st));
  CK(cudaStreamSynchronize(st));  // lk staging buffer is a local
  m->k = 0;
  return SCFTB_OK;
}

// one SCFT iteration of every problem: Y_k = F(X_k), then the Anderson update; asynchronous on `stream`
int scftb_mixer_iterate_device(scftb_mixer *m, void *stream) {
  if (!m) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  scftb_engine *e = m->e;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (e->params_dirty) {
    int rc = upload_params(e);
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->stream));
  }
  const long long stride = (long long)m->R * m->n;
  const size_t slot = (size_t)(m->k % m->R) * m->n;
  int rc = e->diblock_state ? launch_residual_ab(e, m->nprob, m->X + slot, stride, m->Y + slot, stride, m->freeze ? m->done : nullptr, st)
                            : launch_march(e, m->nprob, m->X + slot, stride, m->Y + slot, stride, m->freeze ? m->done : nullptr, st);
  if (rc) return rc;
  AndersonParams A;
  A.n = m->n; A.nprob = m->nprob; A.R = m->R; A.nm = m->nm; A.k = m->k; A.Final = m->Final;
  A.tol = m->tol; A.lmd = m->lmd; A.freeze = m->freeze;
  A.adm = m->adm;
  A.relax = m->k == 0 ? 0.05 : 1.0 - std::pow(0.95, m->k + 1);   // adm.c:140 (its = 1), adm.c:151 (its = k+1)
  A.X = m->X; A.Y = m->Y; A.xfinal = m->xfinal; A.lk = m->lk; A.err = m->err;
  A.k_restart = m->k_restart; A.done = m->done; A.iters = m->iters;
  anderson_kernel<<<m->nprob, AND_THREADS, m->smem, st>>>(A);
  g_launches++;
  CK(cudaGetLastError());
  m->k++;
  return SCFTB_OK;
}

// per-problem state after the iterations issued so far (synchronises `stream`)
int scftb_mixer_status(scftb_mixer *m, void *stream, int *done, int *iters, double *err) {
  if (!m) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  CK(cudaSetDevice(m->e->cfg.device));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  if (done) CK(cudaMemcpy(done, m->done, sizeof(int) * m->nprob, cudaMemcpyDeviceToHost));
  if (iters) CK(cudaMemcpy(iters, m->iters, sizeof(int) * m->nprob, cudaMemcpyDeviceToHost));
  if (err) CK(cudaMemcpy(err, m->err, sizeof(double) * m->nprob, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

// current fields: the converged X[k] of finished problems, the field whose residual first went NaN for done == 2,
// the latest iterate X[k] of the others
int scftb_mixer_get_x(scftb_mixer *m, void *stream, double *x) {
  if (!m || !x) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  scftb_engine *e = m->e;
  const size_t n = m->n;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  std::vector<int> done(m->nprob);
  CK(cudaMemcpy(done.data(), m->done, sizeof(int) * m->nprob, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy2D(x, sizeof(double) * n, m->X + (size_t)(m->k % m->R) * n, sizeof(double) * m->R * n, sizeof(double) * n,
                  m->nprob, cudaMemcpyDeviceToHost));
  for (int p = 0; p < m->nprob; p++)
    if (done[p] != 0) CK(cudaMemcpy(x + (size_t)p * n, m->xfinal + (size_t)p * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

int scftb_mixer_get_y(scftb_mixer *m, void *stream, int k, double *y) {
  if (!m || !y) return fail(SCFTB_ERR_ARG, "mixer: null argument");
  if (k < 0 || k >= m->k || k < m->k - m->R) return fail(SCFTB_ERR_STATE, "mixer: residual of that iteration is not in the ring");
  const size_t n = m->n;
  CK(cudaSetDevice(m->e->cfg.device));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  CK(cudaMemcpy2D(y, sizeof(double) * n, m->Y + (size_t)(k % m->R) * n, sizeof(double) * m->R * n, sizeof(double) * n,
                  m->nprob, cudaMemcpyDeviceToHost));
  return SCFTB_OK;
}

static int mixer_batch(scftb_mixer *m, scftb_engine *e, int nprob, double *x, int maxIteration, int *iters_out, double *err_out);

int scftb_adm_chen_batch(scftb_engine *e, int nprob, double *x, double tol, int maxIteration, double lmd, int nn,
                         int Final, int *iters_out, double *err_out) {
  if (!e || !x || maxIteration < 0) return fail(SCFTB_ERR_ARG, "adm_chen_batch: bad argument");
  scftb_mixer *m = nullptr;
  int rc = scftb_mixer_create(e, nprob, tol, lmd, nn, Final, &m);
  if (rc) return rc;
  return mixer_batch(m, e, nprob, x, maxIteration, iters_out, err_out);
}

// adm for a batch, everything on the device: x[nprob][N-2] host in/out; at most maxits residual evaluations per problem
// (adm.c MAXITS); iters_out / err_out as scftb_adm_chen_batch; returns 0 when every problem reached adm's TOLF = 1e-10
int scftb_adm_batch(scftb_engine *e, int nprob, double *x, int maxits, int *iters_out, double *err_out) {
  if (!e || !x || maxits < 1) return fail(SCFTB_ERR_ARG, "adm_batch: bad argument");
  scftb_mixer *m = nullptr;
  int rc = scftb_adm_mixer_create(e, nprob, &m);
  if (rc) return rc;
  return mixer_batch(m, e, nprob, x, maxits - 1, iters_out, err_out);
}

static int mixer_batch(scftb_mixer *m, scftb_engine *e, int nprob, double *x, int maxIteration, int *iters_out, double *err_out) {
  int rc;
  rc = scftb_mixer_reset(m, x, 0, nullptr);
  std::vector<int> done(nprob, 0), iters(nprob, 0);
  bool all = false, nanseen = false;
  // the reference evaluates F at k = 0..maxIteration (ADM_chen_C.c:55); poll the flags every few iterations
  for (int k = 0; !rc && k <= maxIteration && !all; k++) {
    rc = scftb_mixer_iterate_device(m, e->stream);
    if (!rc && (k % 8 == 7 || k == maxIteration)) {
      rc = scftb_mixer_status(m, e->stream, done.data(), nullptr, nullptr);
      all = std::all_of(done.begin(), done.end(), [](int d) { return d != 0; });
    }
  }
  if (!rc) rc = scftb_mixer_status(m, e->stream, done.data(), iters.data(), err_out);
  if (!rc) rc = scftb_mixer_get_x(m, e->stream, x);
  int status = SCFTB_OK;
  for (int p = 0; p < nprob; p++) {
    if (done[p] == 2) nanseen = true;
    if (done[p] == 0) { status = SCFTB_ERR_NOCONV; iters[p] = m->k; }
    if (iters_out) iters_out[p] = iters[p];
  }
  scftb_mixer_destroy(m);
  if (rc) return rc;
  if (nanseen) return fail(SCFTB_ERR_NAN, "adm_chen_batch: NaN residual");
  if (status) fail(status, "adm_chen_batch: iteration limit reached for at least one problem");
  return status;
}

}  // extern "C"
