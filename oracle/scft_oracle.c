/*
 * scft_oracle.c — CPU oracle for the SCFT propagator hot path (see scft_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 * Citations are file:line under /root/reference (giantsda/SCFT).
 */
#include "scft_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * General band LU with partial pivoting (stands in for UMFPACK / MATLAB backslash / inv(D):
 * 1D_FEM.c:215 KSPSolve, drivescft.cc:141 A_direct.vmult, simple_FEM_1D_transient.m:80,91).
 * Column-major band storage, ldab = 2*kl+ku+1, A(i,j) at ab[kl+ku+i-j + j*ldab].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n, kl, ku, ldab;
  double *ab;
  int *ipiv;
} band_t;

static band_t *band_new(int n, int kl, int ku) {
  band_t *m = (band_t *)malloc(sizeof(band_t));
  m->n = n; m->kl = kl; m->ku = ku; m->ldab = 2 * kl + ku + 1;
  m->ab = (double *)calloc((size_t)m->ldab * n, sizeof(double));
  m->ipiv = (int *)malloc(sizeof(int) * n);
  return m;
}
static void band_free(band_t *m) { free(m->ab); free(m->ipiv); free(m); }
static void band_set(band_t *m, int i, int j, double v) {
  m->ab[m->kl + m->ku + i - j + (size_t)j * m->ldab] = v;
}
static void band_factor(band_t *m) {
  const int n = m->n, kl = m->kl, ku = m->ku, ld = m->ldab, kv = kl + ku;
  double *ab = m->ab;
  int ju = 0;
  for (int j = 0; j < n; j++) {
    int km = kl < n - 1 - j ? kl : n - 1 - j;
    int jp = 0;
    double big = fabs(ab[kv + (size_t)j * ld]);
    for (int i = 1; i <= km; i++) {
      double v = fabs(ab[kv + i + (size_t)j * ld]);
      if (v > big) { big = v; jp = i; }
    }
    m->ipiv[j] = j + jp;
    int cand = j + ku + jp; if (cand > n - 1) cand = n - 1;
    if (cand > ju) ju = cand;
    if (jp != 0)
      for (int c = j; c <= ju; c++) {
        double *p = &ab[kv + jp - (c - j) + (size_t)c * ld], *q = &ab[kv - (c - j) + (size_t)c * ld];
        double t = *p; *p = *q; *q = t;
      }
    double piv = ab[kv + (size_t)j * ld];
    for (int i = 1; i <= km; i++) ab[kv + i + (size_t)j * ld] /= piv;
    for (int c = j + 1; c <= ju; c++) {
      double u = ab[kv - (c - j) + (size_t)c * ld];
      if (u != 0.0)
        for (int i = 1; i <= km; i++)
          ab[kv + i - (c - j) + (size_t)c * ld] -= ab[kv + i + (size_t)j * ld] * u;
    }
  }
}
static void band_solve(const band_t *m, double *b) {
  const int n = m->n, kl = m->kl, ld = m->ldab, kv = m->kl + m->ku;
  const double *ab = m->ab;
  for (int j = 0; j < n; j++) {
    int p = m->ipiv[j];
    if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
    int km = kl < n - 1 - j ? kl : n - 1 - j;
    double bj = b[j];
    for (int i = 1; i <= km; i++) b[j + i] -= ab[kv + i + (size_t)j * ld] * bj;
  }
  for (int j = n - 1; j >= 0; j--) {
    b[j] /= ab[kv + (size_t)j * ld];
    double bj = b[j];
    int lo = j - kv; if (lo < 0) lo = 0;
    for (int i = lo; i < j; i++) b[i] -= ab[kv - (j - i) + (size_t)j * ld] * bj;
  }
}

/* ------------------------------------------------------------------------------------------
 * Romberg integration (romint.c:21-57) with K=5 Neville extrapolation (polint.c:5-42).
 * ------------------------------------------------------------------------------------------ */
#define ORC_ROM_K 5

/* polynomial through (xa[i],ya[i]), i<n, evaluated at x (polint.c:5-42, 0-based restatement) */
static double neville(const double *xa, const double *ya, int n, double x) {
  double c[16], d[16];
  int ns = 0;
  double dif = fabs(x - xa[0]);
  for (int i = 0; i < n; i++) {
    double dift = fabs(x - xa[i]);
    if (dift < dif) { ns = i; dif = dift; }
    c[i] = ya[i]; d[i] = ya[i];
  }
  double y = ya[ns--];
  for (int m = 1; m < n; m++) {
    for (int i = 0; i < n - m; i++) {
      double ho = xa[i] - x, hp = xa[i + m] - x, w = c[i + 1] - d[i];
      double den = ho - hp;
      den = w / den;
      d[i] = hp * den;
      c[i] = ho * den;
    }
    /* polint.c:38 with 1-based ns' = ns+1: 2*ns' < n-m ? c[ns'+1] : d[ns'--] */
    double dy = (2 * (ns + 1) < (n - m)) ? c[ns + 1] : d[ns--];
    y += dy;
  }
  return y;
}

double orc_romint(const double *f, int m, double hh) {
  int M = (int)(log(m * 1.0) / 0.6931471805599453 + 1.5); /* romint.c:28 */
  if (M < ORC_ROM_K) { fprintf(stderr, "orc_romint: m must be >= 16\n"); exit(1); }
  double *s = (double *)malloc(sizeof(double) * (M + 1));
  double *h = (double *)malloc(sizeof(double) * (M + 1));
  h[1] = 1.0;
  s[1] = m * hh * (f[0] + f[m]) / 2;                       /* romint.c:38 */
  int np = 1;
  for (int j = 2; j <= M; j++) {                            /* romint.c:41-48 */
    int ii = m / np;
    double sum = 0;
    for (int i = ii / 2; i < m; i += ii) sum += f[i];
    s[j] = (s[j - 1] + ii * hh * sum) / 2;
    np += np;
    h[j] = h[j - 1] / 4;
  }
  /* romint.c:50: polint(&h[M-K], &s[M-K], K, 0) uses 1-based entries M-K+1..M */
  double ss = neville(&h[M - ORC_ROM_K + 1], &s[M - ORC_ROM_K + 1], ORC_ROM_K, 0.0);
  free(s); free(h);
  return ss;
}

void orc_romberg_weights(int m, double hh, double *w) {
  /* romint is linear in f: ss = sum_k c_k T_k over the last K trapezoid levels, and the
   * Neville coefficients c_k depend only on the abscissae h_k = 4^-(k-1). */
  int M = (int)(log(m * 1.0) / 0.6931471805599453 + 1.5);
  double h[64], ck[ORC_ROM_K];
  h[1] = 1.0;
  for (int j = 2; j <= M; j++) h[j] = h[j - 1] / 4;
  for (int k = 0; k < ORC_ROM_K; k++) {
    double e[ORC_ROM_K] = {0, 0, 0, 0, 0};
    e[k] = 1.0;
    ck[k] = neville(&h[M - ORC_ROM_K + 1], e, ORC_ROM_K, 0.0);
  }
  for (int i = 0; i <= m; i++) w[i] = 0.0;
  for (int k = 0; k < ORC_ROM_K; k++) {
    int level = M - ORC_ROM_K + 1 + k;          /* trapezoid with 2^(level-1) intervals */
    int stride = m >> (level - 1);
    double hl = hh * stride;
    for (int i = 0; i <= m; i += stride)
      w[i] += ck[k] * hl * ((i == 0 || i == m) ? 0.5 : 1.0);
  }
}

/* ------------------------------------------------------------------------------------------
 * Target density (scft.cc:188-215).
 * ------------------------------------------------------------------------------------------ */
static double f0_point(double x, double tau) {
  double e = exp(4 * tau * x / (tau * tau - x * x));
  double v = pow(e - 1, 2) / pow(e + 1, 2);
  if (isnan(v)) v = 1.0;
  return v;
}

void orc_f0_given(int N, const double *x, double tau, double *f0) {
  for (int i = 0; i < N; i++) f0[i] = 1.0;
  for (int i = 0; i < N; i++) {
    if (x[i] <= tau) {
      f0[i] = f0_point(x[i], tau);
      f0[N - i - 1] = f0[i];
    } else
      break;
  }
}

double orc_f0bar(double tau, double L) { /* testFiBar.cc:19-50 */
  int N = (1 << 16) + 1;
  double *x = (double *)malloc(sizeof(double) * N), *f = (double *)malloc(sizeof(double) * N);
  for (int i = 0; i < N; i++) x[i] = L / (N - 1) * i;
  orc_f0_given(N, x, tau, f);
  double r = orc_romint(f, N - 1, L / (N - 1)) / L;
  free(x); free(f);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Cubic spline (spline_chen.c:12-106); the reference solves the same rows with dense gaussj.
 * ------------------------------------------------------------------------------------------ */
void orc_spline(const double *x, const double *y, const double *xp, double *yp, int Nx, int Nxp,
                int mode, double bc) {
  band_t *A = band_new(Nx, 2, 2);
  double *B = (double *)calloc(Nx, sizeof(double));
  for (int i = 1; i < Nx - 1; i++) {                       /* spline_chen.c:32-39 (1-based i+1) */
    band_set(A, i, i - 1, (x[i] - x[i - 1]) / 6.);
    band_set(A, i, i, (x[i + 1] - x[i - 1]) / 3.);
    band_set(A, i, i + 1, (x[i + 1] - x[i]) / 6.);
    B[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
  }
  if (mode == ORC_SPLINE_NOTAKNOT) {                        /* spline_chen.c:41-57 */
    if (Nx <= 3) { fprintf(stderr, "orc_spline: not-a-knot needs > 3 points\n"); exit(1); }
    band_set(A, 0, 0, 1 / (x[1] - x[0]));
    band_set(A, 0, 1, -1 / (x[1] - x[0]) - 1 / (x[2] - x[1]));
    band_set(A, 0, 2, 1 / (x[2] - x[1]));
    band_set(A, Nx - 1, Nx - 3, 1 / (x[Nx - 2] - x[Nx - 3]));
    band_set(A, Nx - 1, Nx - 2, -1 / (x[Nx - 2] - x[Nx - 3]) - 1 / (x[Nx - 1] - x[Nx - 2]));
    band_set(A, Nx - 1, Nx - 1, 1 / (x[Nx - 1] - x[Nx - 2]));
  } else {                                                  /* spline_chen.c:58-64 */
    double m = (mode == ORC_SPLINE_NATURAL) ? 0.0 : bc;
    band_set(A, 0, 0, 1.);
    band_set(A, Nx - 1, Nx - 1, 1.);
    B[0] = m; B[Nx - 1] = m;
  }
  band_factor(A);
  band_solve(A, B);
  for (int i = 0; i < Nxp; i++) {                           /* spline_chen.c:76-100 */
    int klo = 0, khi = Nx - 1;
    while (khi - klo > 1) {
      int k = (khi + klo) >> 1;
      if (x[k] > xp[i]) khi = k; else klo = k;
    }
    double h = x[khi] - x[klo];
    double a = (x[khi] - xp[i]) / h, b = (xp[i] - x[klo]) / h;
    yp[i] = a * y[klo] + b * y[khi] + ((a * a * a - a) * B[klo] + (b * b * b - b) * B[khi]) * (h * h) / 6.0;
  }
  band_free(A); free(B);
}

static void mesh_coords(const orc_config *cfg, double *x) {
  if (cfg->x) memcpy(x, cfg->x, sizeof(double) * cfg->N);
  else for (int i = 0; i < cfg->N; i++) x[i] = cfg->L * i / (cfg->N - 1); /* ~ subdivided_hyper_rectangle, drivescft.cc:94 */
}

void orc_eta_full(int N, const double *x, const double *eta_mid, double *eta_full) {
  /* scft.cc:456-475: knots = interior nodes, evaluation points = every node */
  orc_spline(x + 1, eta_mid, x, eta_full, N - 2, N, ORC_SPLINE_NATURAL, 0.0);
}

/* ------------------------------------------------------------------------------------------
 * The residual evaluation.
 * ------------------------------------------------------------------------------------------ */
int orc_residual(const orc_config *cfg, const double *eta, const double *f0_given, double *out_mid,
                 double *phi_out, double *q_hist_out, double *Q_out) {
  const int N = cfg->N, n = cfg->nsteps, ni = N - 2;
  const double dt = 1. / n; /* time_step = 1/(total_time_step-1), scft.cc:29; 1D_FEM.c:61 */
  double *x = (double *)malloc(sizeof(double) * N);
  mesh_coords(cfg, x);
  const int uniform = (cfg->x == NULL);
  const double hU = cfg->L / (N - 1); /* 1D_FEM.c:61 */

  /* tridiagonal rows of A (mass), B (stiffness), C (eta-weighted mass), interior rows only.
   * Dirichlet rows/columns drop out: q(0)=q(L)=0 for all s (1D_FEM.c:179-182,213-214;
   * scft.cc:599-606,660-665 — constrained dofs decouple under distribute_local_to_global). */
  double *Al = calloc(N, 8), *Ad = calloc(N, 8), *Au = calloc(N, 8);
  double *Dl = calloc(N, 8), *Dd = calloc(N, 8), *Du = calloc(N, 8); /* D = B + C */
  for (int i = 1; i <= N - 2; i++) {
    double a1 = uniform ? hU : x[i] - x[i - 1], a2 = uniform ? hU : x[i + 1] - x[i];
    double bl, bd, bu, cl, cd, cu;
    if (uniform) { /* 1D_FEM.c:95-98 */
      Al[i] = hU / 6; Ad[i] = 2. / 3 * hU; Au[i] = hU / 6;
      bl = -1 / hU; bd = 2. / hU; bu = -1 / hU;
    } else {       /* simple_FEM_1D_transient.m:35-57 */
      Al[i] = a1 / 6; Ad[i] = a1 / 3 + a2 / 3; Au[i] = a2 / 6;
      bl = -1 / a1; bd = 1 / a1 + 1 / a2; bu = -1 / a2;
    }
    if (cfg->scheme == ORC_IE_ROWSCALE) { /* 1D_FEM.c:104-105; simple_FEM_1D_transient.m:66-70 */
      cl = Al[i] * eta[i]; cd = Ad[i] * eta[i]; cu = Au[i] * eta[i];
    } else { /* (eta_h phi_i, phi_j) with eta_h piecewise linear; 2-pt Gauss is exact (scft.cc:653-655) */
      cl = a1 * (eta[i - 1] + eta[i]) / 12;
      cu = a2 * (eta[i] + eta[i + 1]) / 12;
      cd = a1 * (eta[i - 1] + 3 * eta[i]) / 12 + a2 * (3 * eta[i] + eta[i + 1]) / 12;
    }
    Dl[i] = bl + cl; Dd[i] = bd + cd; Du[i] = bu + cu;
  }

  double *hist = (double *)calloc((size_t)N * (n + 1), sizeof(double));
#define H(i, j) hist[(size_t)(i) * (n + 1) + (j)]
  double *q = calloc(N, 8), *rhs = calloc(2 * (size_t)N, 8);
  for (int i = 1; i <= N - 2; i++) { q[i] = 1.0; H(i, 0) = 1.0; } /* drivescft.cc:120-127 */

  if (cfg->scheme == ORC_IRK4_CONSISTENT) {
    /* scft.cc:671-693, unknowns interleaved (k1_i, k2_i) so the block matrix is banded */
    const double c01 = (1. / 4. - sqrt(3) / 6.) * dt, c10 = (1. / 4 + sqrt(3) / 6.) * dt;
    band_t *S = band_new(2 * ni, 3, 3);
    for (int i = 1; i <= N - 2; i++) {
      int r = 2 * (i - 1);
      for (int d = -1; d <= 1; d++) {
        int j = i + d;
        if (j < 1 || j > N - 2) continue;
        double a = d < 0 ? Al[i] : (d == 0 ? Ad[i] : Au[i]);
        double dd = d < 0 ? Dl[i] : (d == 0 ? Dd[i] : Du[i]);
        int c = 2 * (j - 1);
        band_set(S, r, c, a + dt / 4 * dd);
        band_set(S, r, c + 1, c01 * dd);
        band_set(S, r + 1, c, c10 * dd);
        band_set(S, r + 1, c + 1, a + dt / 4 * dd);
      }
    }
    band_factor(S);
    for (int step = 1; step <= n; step++) { /* drivescft.cc:130-152 */
      for (int i = 1; i <= N - 2; i++) {
        double t = Dl[i] * q[i - 1] + Dd[i] * q[i] + Du[i] * q[i + 1];
        rhs[2 * (i - 1)] = -t; rhs[2 * (i - 1) + 1] = -t;
      }
      band_solve(S, rhs);
      for (int i = 1; i <= N - 2; i++) {
        q[i] = q[i] + 0.5 * dt * (rhs[2 * (i - 1)] + rhs[2 * (i - 1) + 1]);
        H(i, step) = q[i];
      }
    }
    band_free(S);
  } else {
    /* D_IE = A + dt*B + dt*C (1D_FEM.c:177-178; simple_FEM_1D_transient.m:74) */
    band_t *S = band_new(ni, 1, 1);
    for (int i = 1; i <= N - 2; i++) {
      int r = i - 1;
      if (i > 1) band_set(S, r, r - 1, Al[i] + dt * Dl[i]);
      band_set(S, r, r, Ad[i] + dt * Dd[i]);
      if (i < N - 2) band_set(S, r, r + 1, Au[i] + dt * Du[i]);
    }
    band_factor(S);
    for (int step = 1; step <= n; step++) { /* 1D_FEM.c:208-228 */
      for (int i = 1; i <= N - 2; i++) rhs[i - 1] = Al[i] * q[i - 1] + Ad[i] * q[i] + Au[i] * q[i + 1];
      band_solve(S, rhs);
      for (int i = 1; i <= N - 2; i++) { q[i] = rhs[i - 1]; H(i, step) = q[i]; }
    }
    band_free(S);
  }

  /* density quadrature (drivescft.cc:184-193) */
  double *v = (double *)malloc(sizeof(double) * (n + 1));
  double *phi = (double *)calloc(N, sizeof(double));
  for (int i = 0; i < N; i++) {
    for (int j = 0; j <= n; j++) v[j] = H(i, j) * H(i, n - j);
    if (cfg->quadrature == ORC_QUAD_ROMBERG)
      phi[i] = orc_romint(v, n, 1. / n);
    else { /* simple_FEM_1D_transient.m:120-124 */
      double s = 0;
      for (int j = 0; j < n; j++) s = s + 0.5 * (v[j] + v[j + 1]) * dt;
      phi[i] = s;
    }
  }
  if (out_mid)
    for (int i = 1; i <= N - 2; i++) out_mid[i - 1] = cfg->sign * (f0_given[i] - phi[i]); /* drivescft.cc:210-213 */
  if (phi_out) memcpy(phi_out, phi, sizeof(double) * N);
  if (q_hist_out) memcpy(q_hist_out, hist, sizeof(double) * (size_t)N * (n + 1));
  if (Q_out) {
    double s = 0;
    for (int i = 1; i <= N - 2; i++) {
      double a1 = uniform ? hU : x[i] - x[i - 1], a2 = uniform ? hU : x[i + 1] - x[i];
      s += 0.5 * (a1 + a2) * q[i];
    }
    *Q_out = s / (uniform ? cfg->L : (x[N - 1] - x[0]));
  }
#undef H
  free(x); free(Al); free(Ad); free(Au); free(Dl); free(Dd); free(Du);
  free(hist); free(q); free(rhs); free(v); free(phi);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Two-species (AB diblock) extension, SURVEY.md section 8(f)-4.  NOT in the reference: the
 * reference melt is one species and uses q+(x,s) = q(x,1-s) (drivescft.cc:189-190).  Here the
 * chain has an A block (contour steps 1..jf, field eta_A) and a B block (steps jf+1..n, field
 * eta_B); q is marched from the A end, q+ from the B end, each with the implicit-Euler step of
 * orc_residual, and
 *     phi_A(x) = int_0^f q(x,s) q+(x,1-s) ds,   phi_B(x) = int_f^1 q(x,s) q+(x,1-s) ds
 * with the reference's quadrature on each block (Romberg when the block has 2^k >= 16 steps,
 * else the trapezoid rule of simple_FEM_1D_transient.m:120-124).  Residual, same gauge as the
 * reference (no 1/Q):   out[0..ni)   = sign*(phi_0 - phi_A - phi_B)         incompressibility
 *                       out[ni..2ni) = eta_A - eta_B - chiN*(phi_B - phi_A)  exchange
 * With eta_A == eta_B both sweeps coincide and phi_A + phi_B is the one-sweep density.
 * PARITY PINNED BY ORACLE ONLY.
 * ------------------------------------------------------------------------------------------ */
static band_t *ie_system(const orc_config *cfg, const double *x, const double *eta, double *Al, double *Ad,
                         double *Au) {
  const int N = cfg->N, n = cfg->nsteps, ni = N - 2;
  const double dt = 1. / n, hU = cfg->L / (N - 1);
  const int uniform = (cfg->x == NULL);
  band_t *S = band_new(ni, 1, 1);
  for (int i = 1; i <= N - 2; i++) {
    double a1 = uniform ? hU : x[i] - x[i - 1], a2 = uniform ? hU : x[i + 1] - x[i];
    double bl, bd, bu, cl, cd, cu;
    if (uniform) { Al[i] = hU / 6; Ad[i] = 2. / 3 * hU; Au[i] = hU / 6; bl = -1 / hU; bd = 2. / hU; bu = -1 / hU; }
    else { Al[i] = a1 / 6; Ad[i] = a1 / 3 + a2 / 3; Au[i] = a2 / 6; bl = -1 / a1; bd = 1 / a1 + 1 / a2; bu = -1 / a2; }
    if (cfg->scheme == ORC_IE_ROWSCALE) { cl = Al[i] * eta[i]; cd = Ad[i] * eta[i]; cu = Au[i] * eta[i]; }
    else {
      cl = a1 * (eta[i - 1] + eta[i]) / 12;
      cu = a2 * (eta[i] + eta[i + 1]) / 12;
      cd = a1 * (eta[i - 1] + 3 * eta[i]) / 12 + a2 * (3 * eta[i] + eta[i + 1]) / 12;
    }
    int r = i - 1;
    if (i > 1) band_set(S, r, r - 1, Al[i] + dt * (bl + cl));
    band_set(S, r, r, Ad[i] + dt * (bd + cd));
    if (i < N - 2) band_set(S, r, r + 1, Au[i] + dt * (bu + cu));
  }
  band_factor(S);
  return S;
}

static double quad_block(const orc_config *cfg, const double *v, int m, double h) {
  if (cfg->quadrature == ORC_QUAD_ROMBERG && m >= 16 && (m & (m - 1)) == 0) return orc_romint(v, m, h);
  double s = 0;
  for (int j = 0; j < m; j++) s = s + 0.5 * (v[j] + v[j + 1]) * h;
  return s;
}

int orc_residual_ab(const orc_config *cfg, const double *etaA, const double *etaB, int jf, double chiN,
                    const double *f0_given, double *out, double *phiA_out, double *phiB_out, double *Q_out) {
  const int N = cfg->N, n = cfg->nsteps, ni = N - 2;
  if (cfg->scheme == ORC_IRK4_CONSISTENT || jf < 1 || jf >= n) return 1;
  double *x = (double *)malloc(sizeof(double) * N);
  mesh_coords(cfg, x);
  const int uniform = (cfg->x == NULL);
  const double hU = cfg->L / (N - 1);
  double *Al = calloc(N, 8), *Ad = calloc(N, 8), *Au = calloc(N, 8);
  band_t *SA = ie_system(cfg, x, etaA, Al, Ad, Au);
  band_t *SB = ie_system(cfg, x, etaB, Al, Ad, Au);   /* the mass matrix does not depend on the field */
  double *hq = (double *)calloc((size_t)N * (n + 1), sizeof(double));
  double *hd = (double *)calloc((size_t)N * (n + 1), sizeof(double));
  double *q = calloc(N, 8), *rhs = calloc(N, 8);
  double Q = 0;
  for (int sweep = 0; sweep < 2; sweep++) {
    double *H = sweep == 0 ? hq : hd;
    for (int i = 0; i < N; i++) q[i] = (i >= 1 && i <= N - 2) ? 1.0 : 0.0;
    for (int i = 0; i < N; i++) H[(size_t)i * (n + 1)] = q[i];
    for (int step = 1; step <= n; step++) {
      /* forward: A block first; backward: B block first */
      const band_t *S = (sweep == 0) ? (step <= jf ? SA : SB) : (step <= n - jf ? SB : SA);
      for (int i = 1; i <= N - 2; i++) rhs[i - 1] = Al[i] * q[i - 1] + Ad[i] * q[i] + Au[i] * q[i + 1];
      band_solve(S, rhs);
      for (int i = 1; i <= N - 2; i++) { q[i] = rhs[i - 1]; H[(size_t)i * (n + 1) + step] = q[i]; }
    }
    if (sweep == 0) {
      double s = 0;
      for (int i = 1; i <= N - 2; i++) {
        double a1 = uniform ? hU : x[i] - x[i - 1], a2 = uniform ? hU : x[i + 1] - x[i];
        s += 0.5 * (a1 + a2) * q[i];
      }
      Q = s / (uniform ? cfg->L : (x[N - 1] - x[0]));
    }
  }
  double *v = (double *)malloc(sizeof(double) * (n + 1));
  double *pa = calloc(N, 8), *pb = calloc(N, 8);
  for (int i = 0; i < N; i++) {
    for (int j = 0; j <= n; j++) v[j] = hq[(size_t)i * (n + 1) + j] * hd[(size_t)i * (n + 1) + (n - j)];
    pa[i] = quad_block(cfg, v, jf, 1. / n);
    pb[i] = quad_block(cfg, v + jf, n - jf, 1. / n);
  }
  if (out)
    for (int i = 1; i <= N - 2; i++) {
      out[i - 1] = cfg->sign * (f0_given[i] - pa[i] - pb[i]);
      out[ni + i - 1] = etaA[i] - etaB[i] - chiN * (pb[i] - pa[i]);
    }
  if (phiA_out) memcpy(phiA_out, pa, sizeof(double) * N);
  if (phiB_out) memcpy(phiB_out, pb, sizeof(double) * N);
  if (Q_out) *Q_out = Q;
  band_free(SA); band_free(SB);
  free(x); free(Al); free(Ad); free(Au); free(hq); free(hd); free(q); free(rhs); free(v); free(pa); free(pb);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Free energy (scft.cc:271-291 resample + :404-450 Romberg).
 * ------------------------------------------------------------------------------------------ */
double orc_free_energy(int N, const double *x, const double *eta, double tau, double L, double f0bar,
                       int nplot) {
  double *xp = (double *)malloc(sizeof(double) * nplot), *g = (double *)malloc(sizeof(double) * nplot);
  double *f0 = (double *)malloc(sizeof(double) * nplot);
  for (int i = 0; i < nplot; i++) xp[i] = L * i / (nplot - 1); /* scft.cc:275 */
  orc_f0_given(nplot, xp, tau, f0);                            /* scft.cc:415-436 */
  int k = 0;
  for (int i = 0; i < nplot; i++) { /* FEFieldFunction on Q1: piecewise-linear in x (scft.cc:280-281) */
    while (k < N - 2 && xp[i] > x[k + 1]) k++;
    double t = (xp[i] - x[k]) / (x[k + 1] - x[k]);
    g[i] = ((1 - t) * eta[k] + t * eta[k + 1]) * f0[i];       /* scft.cc:438-441 */
  }
  double I = orc_romint(g, nplot - 1, L / (nplot - 1));         /* scft.cc:443-444 */
  free(xp); free(g); free(f0);
  return (I / f0bar / L + log(f0bar)) / (-1000.);              /* scft.cc:446-447 */
}

/* ------------------------------------------------------------------------------------------
 * Gauss-Jordan, full pivoting (gaussj.c).  Tie-breaking (>=) keeps the LAST largest entry.
 * ------------------------------------------------------------------------------------------ */
int orc_gaussj(double *a, int n, double *b, int m, int variant) {
  int *indxc = malloc(sizeof(int) * n), *indxr = malloc(sizeof(int) * n), *ipiv = calloc(n, sizeof(int));
  int irow = 0, icol = 0, rc = 0;
#define A_(i, j) a[(size_t)(i) * n + (j)]
#define B_(i, j) b[(size_t)(i) * m + (j)]
  for (int i = 0; i < n; i++) {
    double big = 0.0;
    for (int j = 0; j < n; j++)
      if (ipiv[j] != 1)
        for (int k = 0; k < n; k++)
          if (ipiv[k] == 0 && fabs(A_(j, k)) >= big) { big = fabs(A_(j, k)); irow = j; icol = k; }
    ++ipiv[icol];
    if (irow != icol) {
      for (int l = 0; l < n; l++) { double t = A_(irow, l); A_(irow, l) = A_(icol, l); A_(icol, l) = t; }
      for (int l = 0; l < m; l++) { double t = B_(irow, l); B_(irow, l) = B_(icol, l); B_(icol, l) = t; }
    }
    indxr[i] = irow; indxc[i] = icol;
    if (A_(icol, icol) == 0.0) {
      if (variant == 0) { rc = 1; goto done; }       /* DEALII_SCFT/src/gaussj.c:46-50 */
      A_(icol, icol) = A_(icol, icol) + 1e-18;       /* gaussj.c:38 */
    }
    double pivinv = 1.0 / A_(icol, icol);
    A_(icol, icol) = 1.0;
    for (int l = 0; l < n; l++) A_(icol, l) *= pivinv;
    for (int l = 0; l < m; l++) B_(icol, l) *= pivinv;
    for (int ll = 0; ll < n; ll++)
      if (ll != icol) {
        double dum = A_(ll, icol);
        A_(ll, icol) = 0.0;
        for (int l = 0; l < n; l++) A_(ll, l) -= A_(icol, l) * dum;
        for (int l = 0; l < m; l++) B_(ll, l) -= B_(icol, l) * dum;
      }
  }
  for (int l = n - 1; l >= 0; l--)
    if (indxr[l] != indxc[l])
      for (int k = 0; k < n; k++) { double t = A_(k, indxr[l]); A_(k, indxr[l]) = A_(k, indxc[l]); A_(k, indxc[l]) = t; }
done:
#undef A_
#undef B_
  free(indxc); free(indxr); free(ipiv);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * adm_chen (ADM_chen_C.c:18-147): Anderson mixing for F(x)=0 with full X/Y history.
 * ------------------------------------------------------------------------------------------ */
int orc_adm_chen(orc_func f, double *x_old, double tol, int maxIteration, int n, double lmd, int nn,
                 int Final, double *trace, int *iters) {
  int k = 0, k_restart = 0;
  double lk = lmd;
  int nm = nn < n ? nn : n;
  double *U = malloc(sizeof(double) * nm * nm), *V = malloc(sizeof(double) * nm);
  double *X = malloc(sizeof(double) * (size_t)(maxIteration + 2) * n);
  double *Y = malloc(sizeof(double) * (size_t)(maxIteration + 2) * n);
#define X_(r, i) X[(size_t)(r) * n + (i)]
#define Y_(r, i) Y[(size_t)(r) * n + (i)]
  for (int i = 0; i < n; i++) X_(0, i) = x_old[i];
  int rc = 1;
  /* the reference's loop condition reads the outer err (9.9e99, shadowed at :58) => k bound only */
  while (k <= maxIteration) {
    f(n, &X_(k, 0), &Y_(k, 0));
    double err = 0.;
    for (int i = 0; i < n; i++) {
      if (isnan(Y_(k, i))) { rc = 2; goto out; } /* reference exit(1)s here (ADM_chen_C.c:61-66) */
      if (fabs(Y_(k, i)) >= err) err = fabs(Y_(k, i));
    }
    if (trace) trace[k] = err;
    if (err < tol) {
      for (int i = 0; i < n; i++) x_old[i] = X_(k, i);
      rc = 0;
      goto out;
    }
    int m;
    for (;;) { /* restart: label, ADM_chen_C.c:86-112 */
      m = nm < k - k_restart ? nm : k - k_restart;
      for (int i = 0; i < m; i++) {
        for (int j = 0; j < m; j++) {
          double s = 0.;
          for (int t = 0; t < n; t++) s += (Y_(k, t) - Y_(k - i - 1, t)) * (Y_(k, t) - Y_(k - j - 1, t));
          U[i * m + j] = s;
        }
        double s = 0.;
        for (int t = 0; t < n; t++) s += (Y_(k, t) - Y_(k - i - 1, t)) * Y_(k, t);
        V[i] = s;
      }
      if (orc_gaussj(U, m, V, 1, 0) == 1) { k_restart = k; lk = lmd; continue; }
      break;
    }
    for (int i = 0; i < n; i++) { /* ADM_chen_C.c:114-123 */
      double cx = 0., cd = 0.;
      for (int j = 0; j < m; j++) {
        cx += V[j] * (X_(k - j - 1, i) - X_(k, i));
        cd += V[j] * (Y_(k - j - 1, i) - Y_(k, i));
      }
      X_(k + 1, i) = X_(k, i) + cx + (1 - lk) * (Y_(k, i) + cd);
    }
    if (err < 0.03 && k > 100) lk *= lmd;
    if (!Final && lk < 1e-5) lk = lmd;
    if (Final && lk < 1e-15) lk = lmd;
    k++;
  }
  for (int i = 0; i < n; i++) x_old[i] = X_(k, i); /* ADM_chen_C.c:140-141 */
out:
  if (iters) *iters = k;
#undef X_
#undef Y_
  free(U); free(V); free(X); free(Y);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * adm (adm.c:24-313): Anderson mixing for x=f(x), ring of NRMAX+1 slots, lambda=1-0.95^its.
 * ------------------------------------------------------------------------------------------ */
#define ORC_NRMAX 10
int orc_adm(orc_func f, double *x, int n, int maxits, double *trace, int *iters) {
  const double TOLF = 1e-10; /* adm.c:29 */
  const int R = ORC_NRMAX + 1;
  double *xnew = malloc(8 * (size_t)n), *xh = malloc(8 * (size_t)n * R), *dh = malloc(8 * (size_t)n * R);
  double u[ORC_NRMAX * ORC_NRMAX], b[ORC_NRMAX];
  int its = 1, nc, rc = 1;
  double lambda = 0.05, err;
  f(n, x, xnew); /* adm.c:105 */
  for (int i = 0; i < n; i++) { xh[i] = x[i]; dh[i] = xnew[i] - x[i]; }
  nc = 1;
  err = 0;
  for (int i = 0; i < n; i++) { double t = fabs(dh[i]); err = err > t ? err : t; }
  if (trace) trace[0] = err;
  if (err < TOLF) { rc = 0; goto out; }
  for (int i = 0; i < n; i++) x[i] = xh[i] + lambda * dh[i]; /* adm.c:145-146 */
  for (its = 2; its <= maxits; ++its) {
    int nr = its - 1 < ORC_NRMAX ? its - 1 : ORC_NRMAX;
    lambda = 1.0 - pow(0.95, its);
    f(n, x, xnew);
    if (nc == R) nc = 0;
    double *dc = dh + (size_t)n * nc, *xc = xh + (size_t)n * nc;
    for (int i = 0; i < n; i++) { dc[i] = xnew[i] - x[i]; xc[i] = x[i]; }
    int cur = nc;
    ++nc;
    err = 0;
    for (int i = 0; i < n; i++) { double t = fabs(dc[i]); err = err > t ? err : t; }
    if (trace) trace[its - 1] = err;
    if (err < TOLF) { rc = 0; goto out; }
    /* u[p][q] = <d_cur - d_{p back}, d_cur - d_{q back}>, b[p] = <d_cur - d_{p back}, d_cur> (adm.c:232-276) */
    for (int p = 1; p <= nr; p++) {
      const double *dp = dh + (size_t)n * ((cur - p + R) % R);
      for (int q = p; q <= nr; q++) {
        const double *dq = dh + (size_t)n * ((cur - q + R) % R);
        double s = 0.0;
        for (int i = 0; i < n; i++) s += (dc[i] - dq[i]) * (dc[i] - dp[i]);
        u[(p - 1) * nr + (q - 1)] = u[(q - 1) * nr + (p - 1)] = s;
      }
      double s = 0.0;
      for (int i = 0; i < n; i++) s += (dc[i] - dp[i]) * dc[i];
      b[p - 1] = s;
    }
    orc_gaussj(u, nr, b, 1, 1); /* adm.c:278, root gaussj */
    for (int i = 0; i < n; i++) { /* adm.c:289-306 */
      double t = 0.0, t1 = 0.0;
      for (int p = 1; p <= nr; p++) {
        int s = (cur - p + R) % R;
        t += b[p - 1] * (xh[(size_t)n * s + i] - xc[i]);
        t1 += b[p - 1] * (dh[(size_t)n * s + i] - dc[i]);
      }
      x[i] = xc[i] + t + lambda * (dc[i] + t1);
    }
  }
out:
  if (iters) *iters = its;
  free(xnew); free(xh); free(dh);
  return rc;
}
