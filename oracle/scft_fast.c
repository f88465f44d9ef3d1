/*
 * scft_fast.c — the "fair CPU" baseline of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY: linked into
 * oracle/_build/libscft_oracle.so, called by tests/ (checked against orc_residual) and by bench.py's cpu_baseline and
 * `--impl reference` legs, never by the product.
 *
 * Same discretisation as orc_residual for the implicit-Euler schemes on a uniform mesh
 *   assembly   1D_FEM.c:95-105, scft.cc:643-656          D_IE = A + ds (B + C)   1D_FEM.c:177-178
 *   march      1D_FEM.c:208-228                           quadrature  drivescft.cc:184-193, romint.c:21-57
 * written the way a CPU programmer would after profiling the reference (SURVEY.md section 8d "fair CPU"):
 *   - the constant tridiagonal system is factored ONCE per field (Thomas), every contour step is two sweeps
 *     (the reference calls KSPSolve / UMFPACK per step; the checker oracle uses a pivoting band LU);
 *   - only the half of the history the symmetric quadrature re-reads is kept; Romberg is a weight vector;
 *   - VL independent problems of a sweep are interleaved node by node so the sweeps vectorise (AVX-512/AVX2 clones
 *     selected at load time), POSIX threads take groups of VL problems from a shared counter.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

#include "scft_oracle.h"

#define VL 8

/* one group of VL interleaved problems; arrays are [node][VL].  Returns 0. */
__attribute__((target_clones("avx512f", "avx2", "default")))
static int group_march(int N, int n, int scheme, const double *restrict w /* [n+1] */,
                       const double *restrict Lg /* [VL] */, const double *restrict eta /* [ni][VL] interior */,
                       double *restrict phi /* [ni][VL] */, double *restrict qlast /* [ni][VL] */,
                       double *restrict work /* 6*ni*VL */, double *restrict hist /* (n/2+1)*ni*VL */) {
  const int ni = N - 2;
  const double dt = 1.0 / n;
  double *restrict tl = work, *restrict cp = work + (size_t)ni * VL, *restrict inv = work + 2 * (size_t)ni * VL;
  double *restrict q = work + 3 * (size_t)ni * VL, *restrict d = work + 4 * (size_t)ni * VL;
  double ao[VL], bo[VL];
  for (int v = 0; v < VL; v++) { double h = Lg[v] / (N - 1); ao[v] = h / 6; bo[v] = -1 / h; }
  /* rows of T = A + dt (B + C) and their Thomas factors */
  for (int i = 0; i < ni; i++) {
#pragma GCC ivdep
    for (int v = 0; v < VL; v++) {
      const double a = ao[v], b = bo[v];
      double e0 = eta[(size_t)i * VL + v], cl, cd, cu;
      if (scheme == ORC_IE_ROWSCALE) { cl = a * e0; cd = 4 * a * e0; cu = a * e0; }
      else { /* wall values: natural-spline extrapolation, linear on a uniform mesh (scft.cc:452-490) */
        double em = (i > 0) ? eta[(size_t)(i - 1) * VL + v] : 2 * eta[v] - eta[VL + v];
        double ep = (i < ni - 1) ? eta[(size_t)(i + 1) * VL + v]
                                 : 2 * eta[(size_t)(ni - 1) * VL + v] - eta[(size_t)(ni - 2) * VL + v];
        const double h = 6 * a;
        cl = h * (em + e0) / 12; cu = h * (e0 + ep) / 12; cd = h * (em + 3 * e0) / 12 + h * (3 * e0 + ep) / 12;
      }
      double Tl = a + dt * (b + cl), Td = 4 * a + dt * (-2 * b + cd), Tu = a + dt * (b + cu);
      if (i == 0) Tl = 0.0;
      if (i == ni - 1) Tu = 0.0;
      double den = Td - ((i > 0) ? Tl * cp[(size_t)(i - 1) * VL + v] : 0.0);
      tl[(size_t)i * VL + v] = Tl;
      inv[(size_t)i * VL + v] = 1.0 / den;
      cp[(size_t)i * VL + v] = Tu / den;
    }
  }
  for (size_t k = 0; k < (size_t)ni * VL; k++) { q[k] = 1.0; phi[k] = 0.0; hist[k] = 1.0; }
  for (int j = 1; j <= n; j++) {
    /* b = A q on the fly, forward sweep */
    for (int i = 0; i < ni; i++) {
#pragma GCC ivdep
      for (int v = 0; v < VL; v++) {
        const double qm = (i > 0) ? q[(size_t)(i - 1) * VL + v] : 0.0, qp = (i < ni - 1) ? q[(size_t)(i + 1) * VL + v] : 0.0;
        const double b = ao[v] * (qm + 4 * q[(size_t)i * VL + v] + qp);
        const double dm = (i > 0) ? d[(size_t)(i - 1) * VL + v] : 0.0;
        d[(size_t)i * VL + v] = (b - tl[(size_t)i * VL + v] * dm) * inv[(size_t)i * VL + v];
      }
    }
    /* back substitution */
    for (int i = ni - 1; i >= 0; i--) {
#pragma GCC ivdep
      for (int v = 0; v < VL; v++) {
        const double qn = (i < ni - 1) ? q[(size_t)(i + 1) * VL + v] : 0.0;
        q[(size_t)i * VL + v] = d[(size_t)i * VL + v] - cp[(size_t)i * VL + v] * qn;
      }
    }
    if (2 * j <= n) memcpy(hist + (size_t)j * ni * VL, q, sizeof(double) * ni * VL);
    if (2 * j >= n) { /* sum_j w_j q_j q_{n-j}, w symmetric: pairs (j, n-j) for j > n/2, the middle slice once */
      const double wj = (2 * j > n) ? 2 * w[j] : w[j];
      const double *restrict ho = hist + (size_t)(n - j) * ni * VL;
#pragma GCC ivdep
      for (size_t k = 0; k < (size_t)ni * VL; k++) phi[k] += wj * q[k] * ho[k];
    }
  }
  memcpy(qlast, q, sizeof(double) * ni * VL);
  return 0;
}

/* nprob independent problems (tau_p, L_p, eta_p) on uniform N-node meshes.  eta_mid, out: [nprob][N-2];
 * phi: [nprob][N] or NULL; Q: [nprob] or NULL.  Returns 0, or 1 for unsupported arguments / allocation failure. */
typedef struct {
  int nprob, N, n, scheme, ngroups;
  double sign;
  const double *tau, *L, *eta_mid, *w;
  double *out, *phi_out, *Q_out;
  atomic_int next, fail;
} sweep_job;

static void *sweep_worker(void *arg) {
  sweep_job *J = (sweep_job *)arg;
  const int N = J->N, n = J->n, ni = N - 2, nprob = J->nprob;
  double *work = (double *)malloc(sizeof(double) * 6 * (size_t)ni * VL);
  double *hist = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1) * ni * VL);
  double *eta = (double *)malloc(sizeof(double) * (size_t)ni * VL);
  double *phi = (double *)malloc(sizeof(double) * (size_t)ni * VL);
  double *ql = (double *)malloc(sizeof(double) * (size_t)ni * VL);
  double *f0 = (double *)malloc(sizeof(double) * N), *x = (double *)malloc(sizeof(double) * N);
  if (!work || !hist || !eta || !phi || !ql || !f0 || !x) atomic_store(&J->fail, 1);
  else
    for (;;) {
      const int g = atomic_fetch_add(&J->next, 1);
      if (g >= J->ngroups) break;
      double Lg[VL];
      for (int v = 0; v < VL; v++) {
        const int p = (g * VL + v < nprob) ? g * VL + v : nprob - 1; /* pad the last group with a copy */
        Lg[v] = J->L[p];
        for (int i = 0; i < ni; i++) eta[(size_t)i * VL + v] = J->eta_mid[(size_t)p * ni + i];
      }
      group_march(N, n, J->scheme, J->w, Lg, eta, phi, ql, work, hist);
      for (int v = 0; v < VL && g * VL + v < nprob; v++) {
        const int p = g * VL + v;
        for (int i = 0; i < N; i++) x[i] = J->L[p] * i / (N - 1);
        orc_f0_given(N, x, J->tau[p], f0);
        for (int i = 0; i < ni; i++) J->out[(size_t)p * ni + i] = J->sign * (f0[i + 1] - phi[(size_t)i * VL + v]);
        if (J->phi_out) {
          J->phi_out[(size_t)p * N] = J->phi_out[(size_t)p * N + N - 1] = 0.0;
          for (int i = 0; i < ni; i++) J->phi_out[(size_t)p * N + i + 1] = phi[(size_t)i * VL + v];
        }
        if (J->Q_out) {
          double s = 0.0;
          const double h = J->L[p] / (N - 1);
          for (int i = 0; i < ni; i++) s += h * ql[(size_t)i * VL + v];
          J->Q_out[p] = s / J->L[p];
        }
      }
    }
  free(work); free(hist); free(eta); free(phi); free(ql); free(f0); free(x);
  return NULL;
}

int orc_fast_sweep(int nprob, int N, int nsteps, int scheme, int quadrature, double sign, const double *tau,
                   const double *L, const double *eta_mid, double *out, double *phi_out, double *Q_out, int nthreads) {
  if (scheme != ORC_IE_ROWSCALE && scheme != ORC_IE_CONSISTENT) return 1;
  if (N < 5 || nsteps < 2 || nprob < 1) return 1;
  const int n = nsteps;
  double *w = (double *)malloc(sizeof(double) * (n + 1));
  if (!w) return 1;
  if (quadrature == ORC_QUAD_ROMBERG) {
    if (n < 16 || (n & (n - 1))) { free(w); return 1; } /* romint.c:28-33: m = 2^k >= 16 */
    orc_romberg_weights(n, 1.0 / n, w);
  } else {
    for (int j = 0; j <= n; j++) w[j] = (j == 0 || j == n) ? 0.5 / n : 1.0 / n;
  }
  sweep_job J;
  J.nprob = nprob; J.N = N; J.n = n; J.scheme = scheme; J.ngroups = (nprob + VL - 1) / VL; J.sign = sign;
  J.tau = tau; J.L = L; J.eta_mid = eta_mid; J.w = w; J.out = out; J.phi_out = phi_out; J.Q_out = Q_out;
  atomic_init(&J.next, 0); atomic_init(&J.fail, 0);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > J.ngroups) nthreads = J.ngroups;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], NULL, sweep_worker, &J) == 0) started++;
  sweep_worker(&J);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(w);
  return atomic_load(&J.fail);
}
