// postproc.cu — the callers and data formats on either side of the hot path:
//   * cubic spline with the reference's three end conditions (spline_chen.c:12-106), solved as a
//     tridiagonal system instead of dense gaussj
//   * mesh refinement "every cell cut in x" with not-a-knot transfer of the field
//     (scft.cc:132-169), and the refinement-level bookkeeping of the driver loop (drivescft.cc:291-322)
//   * the reference's result-file format: writer (scft.cc:319-337) and reader (scft_util.cc:13-41),
//     and the .res reader of 1D_FEM.c:322-342
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "engine.h"

using namespace scftb;

namespace {

// second derivatives M[0..Nx-1] of the cubic spline through (x,y)
// mode 0: natural (M0 = M_last = 0; spline_chen with *m = 0), 1: not-a-knot (m == NULL), 2: M0 = M_last = bc
int spline_second_derivatives(const double *x, const double *y, int Nx, int mode, double bc, std::vector<double> &M) {
  M.assign(Nx, 0.0);
  if (Nx < 3) return (mode == 1) ? 1 : 0;
  if (mode == 1 && Nx <= 3) return 1;  // spline_chen.c:44-49
  // interior rows i = 1..Nx-2:  h_{i-1}/6 M_{i-1} + (h_{i-1}+h_i)/3 M_i + h_i/6 M_{i+1} = d_i   (spline_chen.c:32-39)
  const int m = Nx - 2;
  std::vector<double> lo(m), di(m), up(m), rh(m);
  for (int i = 1; i <= m; i++) {
    double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
    lo[i - 1] = h0 / 6.; di[i - 1] = (x[i + 1] - x[i - 1]) / 3.; up[i - 1] = h1 / 6.;
    rh[i - 1] = (y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0;
  }
  double a0 = 0, b0 = 0, aN = 0, bN = 0;  // M_0 = a0 M_1 + b0 M_2 (+ bc), M_last = aN M_{last-1} + bN M_{last-2}
  if (mode == 1) {
    // not-a-knot rows (spline_chen.c:50-56): (M1-M0)/h0 = (M2-M1)/h1
    double h0 = x[1] - x[0], h1 = x[2] - x[1];
    a0 = 1.0 + h0 / h1; b0 = -h0 / h1;
    double g0 = x[Nx - 1] - x[Nx - 2], g1 = x[Nx - 2] - x[Nx - 3];
    aN = 1.0 + g0 / g1; bN = -g0 / g1;
    di[0] += lo[0] * a0; up[0] += lo[0] * b0;
    di[m - 1] += up[m - 1] * aN; lo[m - 1] += up[m - 1] * bN;
  } else {
    double e = (mode == 0) ? 0.0 : bc;
    rh[0] -= lo[0] * e;
    rh[m - 1] -= up[m - 1] * e;
    M[0] = M[Nx - 1] = e;
  }
  // Thomas on the m interior unknowns
  std::vector<double> cp(m), dp(m);
  cp[0] = up[0] / di[0]; dp[0] = rh[0] / di[0];
  for (int i = 1; i < m; i++) {
    double den = di[i] - lo[i] * cp[i - 1];
    cp[i] = up[i] / den;
    dp[i] = (rh[i] - lo[i] * dp[i - 1]) / den;
  }
  M[m] = dp[m - 1];
  for (int i = m - 2; i >= 0; i--) M[i + 1] = dp[i] - cp[i] * M[i + 2];
  if (mode == 1) {
    M[0] = a0 * M[1] + b0 * M[2];
    M[Nx - 1] = aN * M[Nx - 2] + bN * M[Nx - 3];
  }
  return 0;
}

}  // namespace

extern "C" {

int scftb_spline(const double *x, const double *y, const double *xp, double *yp, int Nx, int Nxp, int mode, double bc) {
  if (!x || !y || !xp || !yp || Nx < 2 || Nxp < 0) return fail(SCFTB_ERR_ARG, "spline: bad argument");
  std::vector<double> M;
  if (spline_second_derivatives(x, y, Nx, mode, bc, M))
    return fail(SCFTB_ERR_ARG, "spline: not-a-knot needs more than 3 points (spline_chen.c:44-49)");
  for (int i = 0; i < Nxp; i++) {  // bisection + cubic piece, extrapolating outside the knots (spline_chen.c:76-100)
    int klo = 0, khi = Nx - 1;
    while (khi - klo > 1) {
      int k = (khi + klo) >> 1;
      if (x[k] > xp[i]) khi = k; else klo = k;
    }
    double h = x[khi] - x[klo];
    if (h == 0.0) return fail(SCFTB_ERR_ARG, "spline: x must be increasing");
    double a = (x[khi] - xp[i]) / h, b = (xp[i] - x[klo]) / h;
    yp[i] = a * y[klo] + b * y[khi] + ((a * a * a - a) * M[klo] + (b * b * b - b) * M[khi]) * (h * h) / 6.0;
  }
  return SCFTB_OK;
}

// every cell cut in x: N -> 2N-1 nodes (scft.cc:153-157), field on the interior nodes transferred by a
// not-a-knot spline through the old interior nodes (scft.cc:159-166)
int scftb_refine_mesh(int N, const double *x, const double *eta_mid, double *x_new, double *eta_mid_new) {
  if (N < 6 || !x || !eta_mid || !x_new || !eta_mid_new) return fail(SCFTB_ERR_ARG, "refine: bad argument");
  const int Nn = 2 * N - 1;
  for (int i = 0; i < N; i++) x_new[2 * i] = x[i];
  for (int i = 0; i + 1 < N; i++) x_new[2 * i + 1] = 0.5 * (x[i] + x[i + 1]);
  return scftb_spline(x + 1, eta_mid, x_new + 1, eta_mid_new, N - 2, Nn - 2, 1, 0.0);
}

// gradient-based local bisection of Matlab_files/refine_mesh.m:6-30 (the non-uniform-mesh variant): a cell is cut
// when |d eta / dx| across it is >= factor * median over the cells (the two wall cells, whose outer value is
// unknown (Inf in the prototype), are always cut); the field moves to the new interior nodes by the not-a-knot
// spline (refine_mesh.m:41).  Returns the new node count in *N_new; x_new / eta_mid_new need room for 2N-1 / 2N-3.
int scftb_refine_mesh_adaptive(int N, const double *x, const double *eta_mid, double factor, int *N_new, double *x_new,
                               double *eta_mid_new) {
  if (N < 6 || !x || !eta_mid || !N_new || !x_new || !eta_mid_new) return fail(SCFTB_ERR_ARG, "refine: bad argument");
  const int cells = N - 1;
  std::vector<double> err(cells);
  for (int c = 0; c < cells; c++) {
    if (c == 0 || c == cells - 1) { err[c] = INFINITY; continue; }   // solution = [Inf; x_old; Inf]
    err[c] = std::fabs((eta_mid[c] - eta_mid[c - 1]) / (x[c + 1] - x[c]));
  }
  std::vector<double> sorted(err);
  std::sort(sorted.begin(), sorted.end());
  const double med = (cells % 2) ? sorted[cells / 2] : 0.5 * (sorted[cells / 2 - 1] + sorted[cells / 2]);
  const double threshold = med * factor;
  int n = 0;
  for (int c = 0; c < cells; c++) {
    x_new[n++] = x[c];
    if (err[c] >= threshold) x_new[n++] = 0.5 * (x[c] + x[c + 1]);
  }
  x_new[n++] = x[N - 1];
  *N_new = n;
  return scftb_spline(x + 1, eta_mid, x_new + 1, eta_mid_new, N - 2, n - 2, 1, 0.0);
}

// "N= %d, ERROR= %e" / "mean_field_free_energy, %2.15f" / rows "i,x,eta" (scft.cc:328-335)
int scftb_write_solution(const char *path, int N, double err, double F, const double *x, const double *eta_full) {
  FILE *fp = fopen(path, "w+");
  if (!fp) return fail(SCFTB_ERR_ARG, std::string("cannot create file ") + path);
  fprintf(fp, "N= %d, ", N);
  fprintf(fp, "ERROR= %e \n", err);
  fprintf(fp, "mean_field_free_energy, %2.15f \n", F);
  for (int i = 0; i < N; i++) fprintf(fp, "%d,%2.15f,%2.15f\n", i, x[i], eta_full[i]);
  fclose(fp);
  return SCFTB_OK;
}

// detailedsolution_yita_1D_N=<N>.txt (scft.cc:269-312): the piecewise-linear eta_h (FEFieldFunction on the Q1 mesh,
// scft.cc:280-281) sampled on nplot = 2^18 + 1 equidistant points of [0, L], same three-part layout as the solution file
int scftb_write_detailed_solution(const char *path, int N, double err, double F, const double *x, const double *eta_full, int nplot) {
  if (!path || N < 2 || !x || !eta_full) return fail(SCFTB_ERR_ARG, "write_detailed_solution: bad argument");
  if (nplot <= 1) nplot = (1 << 18) + 1;   // scft.cc:271
  FILE *fp = fopen(path, "w+");
  if (!fp) return fail(SCFTB_ERR_ARG, std::string("cannot create file ") + path);
  fprintf(fp, "N= %d, ", N);
  fprintf(fp, "ERROR= %e \n", err);
  fprintf(fp, "mean_field_free_energy, %2.15f \n", F);
  const double L = x[N - 1];
  int k = 0;
  for (int i = 0; i < nplot; i++) {
    const double xp = L * i / (nplot - 1);
    while (k < N - 2 && xp > x[k + 1]) k++;
    const double t = (xp - x[k]) / (x[k + 1] - x[k]);
    fprintf(fp, "%i,%2.15f,%2.15f\n", i, xp, (1 - t) * eta_full[k] + t * eta_full[k + 1]);
  }
  fclose(fp);
  return SCFTB_OK;
}

// reader of the same format (scft_util.cc:13-41); pass x = eta = NULL to query N only
int scftb_read_solution(const char *path, int *N, double *x, double *eta, int capacity) {
  FILE *fp = fopen(path, "r");
  if (!fp) return fail(SCFTB_ERR_ARG, std::string("cannot open file ") + path);
  char buf[255];
  int n = 0;
  if (!fgets(buf, 255, fp) || sscanf(buf, "N= %d", &n) != 1) { fclose(fp); return fail(SCFTB_ERR_ARG, "bad header"); }
  if (N) *N = n;
  long pos = ftell(fp);
  if (fgets(buf, 255, fp) && !strstr(buf, "mean_field_free_energy")) fseek(fp, pos, SEEK_SET);  // older files lack the line
  if (x && eta) {
    if (capacity < n) { fclose(fp); return fail(SCFTB_ERR_ARG, "buffer too small"); }
    for (int i = 0; i < n; i++) x[i] = eta[i] = 0.0;
    int i; double xv, v;
    while (fgets(buf, 255, fp))
      if (sscanf(buf, "%d ,%lf, %lf", &i, &xv, &v) == 3 && i >= 0 && i < n) { x[i] = xv; eta[i] = v; }
  }
  fclose(fp);
  return SCFTB_OK;
}

// Q. Wang's .res files: skip 9 header lines, rows "x/l phi eta phie phij" (1D_FEM.c:322-342, testFiBar.cc:54-79)
int scftb_read_res(const char *path, int rows, double *xl, double *phi, double *eta) {
  FILE *fp = fopen(path, "r");
  if (!fp) return fail(SCFTB_ERR_ARG, std::string("cannot open file ") + path);
  char buf[255];
  for (int i = 0; i < 9; i++)
    if (!fgets(buf, 255, fp)) { fclose(fp); return fail(SCFTB_ERR_ARG, "short .res file"); }
  int got = 0;
  double a, b, c;
  while (got < rows && fgets(buf, 255, fp))
    if (sscanf(buf, "%lf %lf %lf", &a, &b, &c) == 3) {
      if (xl) xl[got] = a;
      if (phi) phi[got] = b;
      if (eta) eta[got] = c;
      got++;
    }
  fclose(fp);
  return got == rows ? SCFTB_OK : fail(SCFTB_ERR_ARG, ".res file has fewer rows than requested");
}

}  // extern "C"
