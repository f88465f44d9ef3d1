// solvers.cu — host-flow field-update solvers behind the reference's own entry points, and the
// reference-shaped residual callbacks (include/scft_b200.h).
//
//   scftb_callback_nr1 / _c0   <->  SCFT_wrapper (drivescft.cc:218-243), simple_FEM_1D_transient
//                                    as passed to broydn (1D_FEM.c:356)
//   scftb_adm_chen             <->  adm_chen   (ADM_chen_C.c:18-147)
//   scftb_adm                  <->  adm        (adm.c:24-313)
//   scftb_broydn               <->  broydn     (broydn.c:44-292) + fdjac.c, lnsrch.c, qrdcmp.c,
//                                    qrupdt.c, rotate.c, rsolv.c
// These keep the reference's iteration logic (same update formulas, pivot order, tolerances and
// return conventions) so that iterates agree with the reference at equal iteration count; the
// residual evaluations they request run on the GPU.  In scftb_broydn the n columns of the
// finite-difference Jacobian are ONE batched launch instead of n sequential evaluations.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "engine.h"

using namespace scftb;

// the engine behind the reference-shaped callbacks, the role of the global `heat_equation_solver` (drivescft.cc:50).
// Per host thread: threads that drive their own engines (sweeps) each bind their own.
static thread_local scftb_engine *g_bound = nullptr;
extern "C" int scftb_funcerr = 0;

namespace scftb {

// Gauss-Jordan elimination with full pivoting on a (n x n, row-major) and one right-hand side.
// Pivot search order and tie-breaking (>=, last maximum wins) as DEALII_SCFT/src/gaussj.c:20-38;
// nudge = true reproduces the root gaussj.c:38 (zero pivot += 1e-18) used by adm.c:278.
int gauss_jordan(std::vector<double> &a, int n, std::vector<double> &b, bool nudge) {
  std::vector<int> colidx(n), rowidx(n), used(n, 0);
  int prow = 0, pcol = 0;
  for (int it = 0; it < n; it++) {
    double big = 0.0;
    for (int j = 0; j < n; j++) {
      if (used[j] == 1) continue;
      for (int k = 0; k < n; k++)
        if (used[k] == 0 && std::fabs(a[j * n + k]) >= big) { big = std::fabs(a[j * n + k]); prow = j; pcol = k; }
    }
    used[pcol]++;
    if (prow != pcol) {
      for (int l = 0; l < n; l++) std::swap(a[prow * n + l], a[pcol * n + l]);
      std::swap(b[prow], b[pcol]);
    }
    rowidx[it] = prow; colidx[it] = pcol;
    if (a[pcol * n + pcol] == 0.0) {
      if (!nudge) return 1;
      a[pcol * n + pcol] += 1e-18;
    }
    const double pinv = 1.0 / a[pcol * n + pcol];
    a[pcol * n + pcol] = 1.0;
    for (int l = 0; l < n; l++) a[pcol * n + l] *= pinv;
    b[pcol] *= pinv;
    for (int r = 0; r < n; r++) {
      if (r == pcol) continue;
      const double f = a[r * n + pcol];
      a[r * n + pcol] = 0.0;
      for (int l = 0; l < n; l++) a[r * n + l] -= a[pcol * n + l] * f;
      b[r] -= b[pcol] * f;
    }
  }
  for (int l = n - 1; l >= 0; l--)
    if (rowidx[l] != colidx[l])
      for (int k = 0; k < n; k++) std::swap(a[k * n + rowidx[l]], a[k * n + colidx[l]]);
  return 0;
}

}  // namespace scftb

extern "C" {

// ------------------------------------------------------------------------------------------------
// called by scftb_destroy: a destroyed engine must not stay bound
void scftb_unbind_engine(scftb_engine *e) { if (g_bound == e) g_bound = nullptr; }

int scftb_bind_global(scftb_engine *e) {
  g_bound = e;
  scftb_funcerr = 0;
  return SCFTB_OK;
}

void scftb_callback_c0(int n, double *in, double *out) {
  (void)n;
  if (!g_bound || scftb_residual(g_bound, in, out) != SCFTB_OK) scftb_funcerr = 1;
}

void scftb_callback_nr1(int n, double *in, double *out) { scftb_callback_c0(n, in + 1, out + 1); }

// two-species residual of the bound engine (n = 2*(N-2) unknowns: eta_A, eta_B)
void scftb_callback_ab_c0(int n, double *in, double *out) {
  (void)n;
  if (!g_bound || scftb_residual_ab(g_bound, in, out) != SCFTB_OK) scftb_funcerr = 1;
}

void scftb_callback_fixedpoint_c0(int n, double *in, double *out) {
  scftb_callback_c0(n, in, out);
  for (int i = 0; i < n; i++) out[i] += in[i];
}

// ------------------------------------------------------------------------------------------------
int scftb_adm_chen(scftb_func f, double *x_old, double tol, int maxIteration, int n, double lmd, int nn, int Final) {
  if (!f || !x_old || n < 1 || maxIteration < 0) return fail(SCFTB_ERR_ARG, "adm_chen: bad argument");
  const int nm = std::min(nn, n);
  std::vector<double> U((size_t)nm * nm), V(nm);
  std::vector<std::vector<double>> X, Y;  // X[k] guessed field, Y[k] its residual (ADM_chen_C.c:49-51)
  X.emplace_back(x_old, x_old + n);
  double lk = lmd;
  int k = 0, k_restart = 0;
  while (k <= maxIteration) {
    Y.emplace_back(n);
    f(n, X[k].data(), Y[k].data());
    double err = 0.0;
    for (int i = 0; i < n; i++) {
      if (std::isnan(Y[k][i])) return fail(SCFTB_ERR_NAN, "adm_chen: NaN residual");  // ADM_chen_C.c:61-66
      if (std::fabs(Y[k][i]) >= err) err = std::fabs(Y[k][i]);
    }
    if (err < tol) {
      std::copy(X[k].begin(), X[k].end(), x_old);
      return 0;
    }
    int m;
    for (;;) {  // ADM_chen_C.c:86-112
      m = std::min(nm, k - k_restart);
      const std::vector<double> &yk = Y[k];
      for (int i = 0; i < m; i++) {
        const std::vector<double> &yi = Y[k - i - 1];
        for (int j = 0; j < m; j++) {
          const std::vector<double> &yj = Y[k - j - 1];
          double s = 0.0;
          for (int t = 0; t < n; t++) s += (yk[t] - yi[t]) * (yk[t] - yj[t]);
          U[i * m + j] = s;
        }
        double s = 0.0;
        for (int t = 0; t < n; t++) s += (yk[t] - yi[t]) * yk[t];
        V[i] = s;
      }
      std::vector<double> Um(U.begin(), U.begin() + (size_t)m * m), Vm(V.begin(), V.begin() + m);
      if (gauss_jordan(Um, m, Vm, false)) { k_restart = k; lk = lmd; continue; }
      std::copy(Vm.begin(), Vm.end(), V.begin());
      break;
    }
    X.emplace_back(n);
    for (int i = 0; i < n; i++) {  // ADM_chen_C.c:114-123
      double cx = 0.0, cd = 0.0;
      for (int j = 0; j < m; j++) {
        cx += V[j] * (X[k - j - 1][i] - X[k][i]);
        cd += V[j] * (Y[k - j - 1][i] - Y[k][i]);
      }
      X[k + 1][i] = X[k][i] + cx + (1 - lk) * (Y[k][i] + cd);
    }
    if (err < 0.03 && k > 100) lk *= lmd;
    if (!Final && lk < 1e-5) lk = lmd;
    if (Final && lk < 1e-15) lk = lmd;
    // history older than the mixing window is never read again
    if (k - nm - 1 >= 0 && k - nm - 1 >= k_restart) { std::vector<double>().swap(X[k - nm - 1]); std::vector<double>().swap(Y[k - nm - 1]); }
    k++;
  }
  std::copy(X[k].begin(), X[k].end(), x_old);  // ADM_chen_C.c:140-141
  return 1;
}

// ------------------------------------------------------------------------------------------------
int scftb_adm(scftb_func f, double *x, int n, int *check, int maxits) {
  if (!f || !x || n < 1 || !check) return fail(SCFTB_ERR_ARG, "adm: bad argument");
  const int NRMAX = 10, R = NRMAX + 1;   // adm.c:6
  const double TOLF = 1e-10;             // adm.c:29
  std::vector<double> xnew(n), xh((size_t)n * R), dh((size_t)n * R), u, b;
  *check = 1;
  f(n, x, xnew.data());
  for (int i = 0; i < n; i++) { xh[i] = x[i]; dh[i] = xnew[i] - x[i]; }
  int nc = 1;
  double err = 0.0;
  for (int i = 0; i < n; i++) err = std::max(err, std::fabs(dh[i]));
  if (err < TOLF) { *check = 0; return 0; }
  double lambda = 0.05;
  for (int i = 0; i < n; i++) x[i] = xh[i] + lambda * dh[i];
  for (int its = 2; its <= maxits; ++its) {
    const int nr = std::min(its - 1, NRMAX);
    lambda = 1.0 - std::pow(0.95, its);   // adm.c:151
    f(n, x, xnew.data());
    if (nc == R) nc = 0;
    double *dc = &dh[(size_t)n * nc], *xc = &xh[(size_t)n * nc];
    for (int i = 0; i < n; i++) { dc[i] = xnew[i] - x[i]; xc[i] = x[i]; }
    const int cur = nc++;
    err = 0.0;
    for (int i = 0; i < n; i++) err = std::max(err, std::fabs(dc[i]));
    if (err < TOLF) { *check = 0; return 0; }
    u.assign((size_t)nr * nr, 0.0); b.assign(nr, 0.0);
    for (int p = 1; p <= nr; p++) {       // adm.c:232-276, index = steps back in the ring
      const double *dp = &dh[(size_t)n * ((cur - p + R) % R)];
      for (int q = p; q <= nr; q++) {
        const double *dq = &dh[(size_t)n * ((cur - q + R) % R)];
        double s = 0.0;
        for (int i = 0; i < n; i++) s += (dc[i] - dq[i]) * (dc[i] - dp[i]);
        u[(p - 1) * nr + q - 1] = u[(q - 1) * nr + p - 1] = s;
      }
      double s = 0.0;
      for (int i = 0; i < n; i++) s += (dc[i] - dp[i]) * dc[i];
      b[p - 1] = s;
    }
    gauss_jordan(u, nr, b, true);         // adm.c:278 with the root gaussj.c
    for (int i = 0; i < n; i++) {         // adm.c:289-306
      double tx = 0.0, td = 0.0;
      for (int p = 1; p <= nr; p++) {
        const size_t s = (size_t)n * ((cur - p + R) % R);
        tx += b[p - 1] * (xh[s + i] - xc[i]);
        td += b[p - 1] * (dh[s + i] - dc[i]);
      }
      x[i] = xc[i] + tx + lambda * (dc[i] + td);
    }
  }
  return 1;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Broyden's method with QR-updated Jacobian and backtracking line search (broydn.c).
// ------------------------------------------------------------------------------------------------
namespace {

struct Mat {  // dense n x n, row-major
  int n;
  std::vector<double> v;
  explicit Mat(int n_) : n(n_), v((size_t)n_ * n_, 0.0) {}
  double &operator()(int i, int j) { return v[(size_t)i * n + j]; }
};

inline double sgn(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }  // nrutil.h SIGN

// Givens rotation of rows i,i+1 of r and qt (rotate.c:5-34)
void rotate_rows(Mat &r, Mat &qt, int i, double a, double b) {
  const int n = r.n;
  double c, s;
  if (a == 0.0) { c = 0.0; s = (b >= 0.0 ? 1.0 : -1.0); }
  else if (std::fabs(a) > std::fabs(b)) { double f = b / a; c = sgn(1.0 / std::sqrt(1.0 + f * f), a); s = f * c; }
  else { double f = a / b; s = sgn(1.0 / std::sqrt(1.0 + f * f), b); c = f * s; }
  for (int j = i; j < n; j++) { double y = r(i, j), w = r(i + 1, j); r(i, j) = c * y - s * w; r(i + 1, j) = s * y + c * w; }
  for (int j = 0; j < n; j++) { double y = qt(i, j), w = qt(i + 1, j); qt(i, j) = c * y - s * w; qt(i + 1, j) = s * y + c * w; }
}

// rank-one update Q R -> Q (R + u v^T) (qrupdt.c:5-24)
void qr_update(Mat &r, Mat &qt, std::vector<double> &u, const std::vector<double> &v) {
  const int n = r.n;
  int k;
  for (k = n - 1; k >= 0; k--) if (u[k] != 0.0) break;
  if (k < 0) k = 0;
  for (int i = k - 1; i >= 0; i--) {
    rotate_rows(r, qt, i, u[i], -u[i + 1]);
    if (u[i] == 0.0) u[i] = std::fabs(u[i + 1]);
    else if (std::fabs(u[i]) > std::fabs(u[i + 1])) { double q = u[i + 1] / u[i]; u[i] = std::fabs(u[i]) * std::sqrt(1.0 + q * q); }
    else { double q = u[i] / u[i + 1]; u[i] = std::fabs(u[i + 1]) * std::sqrt(1.0 + q * q); }
  }
  for (int j = 0; j < n; j++) r(0, j) += u[0] * v[j];
  for (int i = 0; i < k; i++) rotate_rows(r, qt, i, r(i, i), -r(i + 1, i));
}

// Householder QR of a in place (qrdcmp.c:5-33): R above the diagonal, d = diagonal of R
void qr_decompose(Mat &a, std::vector<double> &c, std::vector<double> &d, int &sing) {
  const int n = a.n;
  sing = 0;
  for (int k = 0; k < n - 1; k++) {
    double scale = 0.0;
    for (int i = k; i < n; i++) scale = std::max(scale, std::fabs(a(i, k)));
    if (scale == 0.0) { sing = 1; c[k] = d[k] = 0.0; continue; }
    for (int i = k; i < n; i++) a(i, k) /= scale;
    double sum = 0.0;
    for (int i = k; i < n; i++) sum += a(i, k) * a(i, k);
    double sigma = sgn(std::sqrt(sum), a(k, k));
    a(k, k) += sigma;
    c[k] = sigma * a(k, k);
    d[k] = -scale * sigma;
    for (int j = k + 1; j < n; j++) {
      double s = 0.0;
      for (int i = k; i < n; i++) s += a(i, k) * a(i, j);
      double tau = s / c[k];
      for (int i = k; i < n; i++) a(i, j) -= tau * a(i, k);
    }
  }
  d[n - 1] = a(n - 1, n - 1);
  if (d[n - 1] == 0.0) sing = 1;
}

struct BroydenState {  // what the reference keeps in caller-owned globals qt, r, d (broydn.c:22-28)
  int n = 0;
  std::vector<double> qt, r, d;
};
thread_local BroydenState g_broyden;   // per host thread, like the binding above

}  // namespace

extern "C" int scftb_broydn(scftb_func vecfunc, double *x, int n, int *check, double *err, int *jc) {
  if (!vecfunc || !x || n < 1 || !check || !err || !jc) return fail(SCFTB_ERR_ARG, "broydn: bad argument");
  const int MAXITS = 400;                       // broydn.c:6
  const double EPS = 1e-14, TOLX = EPS, STPMX = 100.0;  // broydn.c:8-11
  const double TOLF = *err, TOLMIN = TOLF;      // broydn.c:65-66
  Mat r(n), qt(n);
  std::vector<double> c(n), d(n), fvec(n), fvcold(n), g(n), p(n), s(n), t(n), w(n), xold(n);
  if (*jc && g_broyden.n == n) { r.v = g_broyden.r; qt.v = g_broyden.qt; d = g_broyden.d; }
  else *jc = 0;
  auto save = [&]() { g_broyden.n = n; g_broyden.r = r.v; g_broyden.qt = qt.v; g_broyden.d = d; };
  scftb_funcerr = 0;
  auto fmin = [&](const double *xx) {           // fminbrd, broydn.c:30-42
    vecfunc(n, const_cast<double *>(xx), fvec.data());
    if (scftb_funcerr) return 0.0;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += fvec[i] * fvec[i];
    return 0.5 * sum;
  };
  auto maxabs = [&](const std::vector<double> &v) { double m = 0.0; for (double a : v) if (std::fabs(a) > m) m = std::fabs(a); return m; };
  auto bail = [&]() { scftb_funcerr = 0; *check = 1; save(); return 0; };  // broydn.c:286-290

  double f = fmin(x);
  if (scftb_funcerr) return bail();
  *err = maxabs(fvec);
  if (*err < TOLF) { *check = 0; return 0; }
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += x[i] * x[i];
  const double stpmax = STPMX * std::max(std::sqrt(sum), (double)n);
  int restrt = (*jc == 0) ? 1 : 0;

  for (int its = 1; its <= MAXITS; its++) {
    if (restrt) {
      // ---- forward-difference Jacobian (fdjac.c:18-34).  The n perturbed evaluations are
      // independent: when the callback is the engine's own, run them as one device batch.
      const double FD = 1.0e-7;
      std::vector<double> hs(n);
      std::vector<double> xb((size_t)n * n), fb((size_t)n * n);
      for (int j = 0; j < n; j++) {
        double temp = x[j], h = FD * temp;
        if (std::fabs(h) < FD) h = sgn(FD, temp);
        double xp = temp + h;
        hs[j] = xp - temp;
        std::copy(x, x + n, xb.begin() + (size_t)j * n);
        xb[(size_t)j * n + j] = xp;
      }
      bool batched = false;
      if (vecfunc == scftb_callback_c0 && g_bound) {
        batched = true;
        int B = scftb_engine_max_batch(g_bound);
        for (int j0 = 0; j0 < n && batched; j0 += B) {
          int nb = std::min(B, n - j0);
          // every column with the parameters of problem 0 (the problem the callback evaluates), whatever (tau, L, mesh)
          // the other slots of a sweep engine hold; their phi/Q/eta_full are left untouched
          if (residual_batch_shared(g_bound, nb, &xb[(size_t)j0 * n], &fb[(size_t)j0 * n]) != SCFTB_OK) { scftb_funcerr = 1; }
        }
      }
      if (vecfunc == scftb_callback_ab_c0 && g_bound) {   // two-species: the 2(N-2) columns in device batches
        batched = true;
        int B = scftb_engine_max_batch(g_bound);
        for (int j0 = 0; j0 < n && batched; j0 += B) {
          int nb = std::min(B, n - j0);
          if (residual_ab_batch_shared(g_bound, nb, &xb[(size_t)j0 * n], &fb[(size_t)j0 * n]) != SCFTB_OK) { scftb_funcerr = 1; }
        }
      }
      if (!batched)
        for (int j = 0; j < n; j++) {
          vecfunc(n, &xb[(size_t)j * n], &fb[(size_t)j * n]);
          if (scftb_funcerr) break;
        }
      if (scftb_funcerr) return bail();
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) r(i, j) = (fb[(size_t)j * n + i] - fvec[i]) / hs[j];
      int sing;
      qr_decompose(r, c, d, sing);
      if (sing) { save(); *check = 1; return fail(SCFTB_ERR_NOCONV, "singular Jacobian in broydn"); }  // broydn.c:128 nrerror
      for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) qt(i, j) = 0.0; qt(i, i) = 1.0; }
      for (int k = 0; k < n - 1; k++)          // form Q^T explicitly (broydn.c:135-149)
        if (c[k] != 0.0)
          for (int j = 0; j < n; j++) {
            double sm = 0.0;
            for (int i = k; i < n; i++) sm += r(i, k) * qt(i, j);
            sm /= c[k];
            for (int i = k; i < n; i++) qt(i, j) -= sm * r(i, k);
          }
      for (int i = 0; i < n; i++) { r(i, i) = d[i]; for (int j = 0; j < i; j++) r(i, j) = 0.0; }
      *jc = 2;
    } else if (its > 1) {                       // Broyden update (broydn.c:158-199)
      for (int i = 0; i < n; i++) s[i] = x[i] - xold[i];
      for (int i = 0; i < n; i++) { double sm = 0.0; for (int j = i; j < n; j++) sm += r(i, j) * s[j]; t[i] = sm; }
      int skip = 1;
      for (int i = 0; i < n; i++) {
        double sm = 0.0;
        for (int j = 0; j < n; j++) sm += qt(j, i) * t[j];
        w[i] = fvec[i] - fvcold[i] - sm;
        if (std::fabs(w[i]) >= EPS * (std::fabs(fvec[i]) + std::fabs(fvcold[i]))) skip = 0;
        else w[i] = 0.0;
      }
      if (!skip) {
        for (int i = 0; i < n; i++) { double sm = 0.0; for (int j = 0; j < n; j++) sm += qt(i, j) * w[j]; t[i] = sm; }
        double den = 0.0;
        for (int i = 0; i < n; i++) den += s[i] * s[i];
        for (int i = 0; i < n; i++) s[i] /= den;
        qr_update(r, qt, t, s);
        for (int i = 0; i < n; i++) {
          if (r(i, i) == 0.0) { save(); *check = 1; return fail(SCFTB_ERR_NOCONV, "r singular in broydn"); }
          d[i] = r(i, i);
        }
      }
    }
    for (int i = 0; i < n; i++) { double sm = 0.0; for (int j = 0; j < n; j++) sm += qt(i, j) * fvec[j]; p[i] = -sm; }
    for (int i = n - 1; i >= 0; i--) { double sm = 0.0; for (int j = 0; j <= i; j++) sm -= r(j, i) * p[j]; g[i] = sm; }
    for (int i = 0; i < n; i++) { xold[i] = x[i]; fvcold[i] = fvec[i]; }
    const double fold = f;
    // solve R p = rhs (rsolv.c:3-13)
    p[n - 1] /= d[n - 1];
    for (int i = n - 2; i >= 0; i--) { double sm = 0.0; for (int j = i + 1; j < n; j++) sm += r(i, j) * p[j]; p[i] = (p[i] - sm) / d[i]; }
    // ---- line search (lnsrch.c:4-60)
    {
      const double ALF = 1.0e-4, LTOLX = 1.0e-7;
      *check = 0;
      double sm = 0.0;
      for (int i = 0; i < n; i++) sm += p[i] * p[i];
      sm = std::sqrt(sm);
      if (sm > stpmax) for (int i = 0; i < n; i++) p[i] *= stpmax / sm;
      double slope = 0.0;
      for (int i = 0; i < n; i++) slope += g[i] * p[i];
      if (slope >= 0.0) { save(); *check = 1; return fail(SCFTB_ERR_NOCONV, "Roundoff problem in lnsrch"); }  // lnsrch.c:18 nrerror
      double test = 0.0;
      for (int i = 0; i < n; i++) test = std::max(test, std::fabs(p[i]) / std::max(std::fabs(xold[i]), 1.0));
      const double alamin = LTOLX / test;
      double alam = 1.0, alam2 = 0.0, f2 = 0.0, tmplam;
      for (;;) {
        for (int i = 0; i < n; i++) x[i] = xold[i] + alam * p[i];
        f = fmin(x);
        if (alam < alamin) { for (int i = 0; i < n; i++) x[i] = xold[i]; *check = 1; break; }
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
        }
        alam2 = alam; f2 = f;
        alam = std::max(tmplam, 0.1 * alam);
      }
    }
    *err = maxabs(fvec);
    if (*err < TOLF) { *check = 0; *jc = 1; save(); return 0; }
    if (*check) {
      if (restrt) { save(); return 0; }         // failed with a fresh Jacobian: jc = 2 (broydn.c:241-243)
      double test = 0.0, den = std::max(f, 0.5 * n);
      for (int i = 0; i < n; i++) test = std::max(test, std::fabs(g[i]) * std::max(std::fabs(x[i]), 1.0) / den);
      if (test < TOLMIN) { *check = 0; *jc = 1; save(); return 0; }
      restrt = 1;
    } else {
      restrt = 0;
      double test = 0.0;
      for (int i = 0; i < n; i++) test = std::max(test, std::fabs(x[i] - xold[i]) / std::max(std::fabs(x[i]), 1.0));
      if (test < TOLX) { *jc = 1; save(); return 0; }
    }
  }
  return bail();
}
