// ref_shim.cc — C-callable entry points onto the UNMODIFIED reference sources, which the
// oracle Makefile compiles in place from /root/reference (as C++, the way the reference's own
// DEALII_SCFT/CMakeLists.txt:27-36 does) into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.
// No reference source is copied: this file only declares the reference's prototypes
// (NR_chen.h:24-38, nr.h) and forwards to them.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

int adm_chen(void (*f)(int, double*, double*), double* x_old, double tol, int maxIteration, int n,
             double lmd, int nn, bool Final);                                   // NR_chen.h:30-34
void spline_chen(double* x, double* y, double* xp, double* yp, int Nx, int Nxp, double* m);  // NR_chen.h:36-38
double romint(double* f, int m, double hh);                                     // romint.c:21
int gaussj(double** a, int n, double** b, int m);                               // DEALII_SCFT/src/gaussj.c:7
void broydn(double x[], int n, int* check, void (*vecfunc)(int, double[], double[]));  // broydn.c:44
double** dmatrix(long nrl, long nrh, long ncl, long nch);                       // nrutil.c
double* dvector(long nl, long nh);
void free_dmatrix(double** m, long nrl, long nrh, long ncl, long nch);
void free_dvector(double* v, long nl, long nh);

// globals the calling program must own (broydn.c:22-28; drivescft.cc:250-256)
double **qt, **r, *d, err;
int funcerr, jc, PRINT;

extern "C" {

double ref_romint(double* f, int m, double hh) { return romint(f, m, hh); }

void ref_spline_chen(double* x, double* y, double* xp, double* yp, int Nx, int Nxp, int mode, double bc) {
  // mode 0: natural (m=0), 1: not-a-knot (m==NULL), 2: given y''
  double m = (mode == 0) ? 0.0 : bc;
  spline_chen(x, y, xp, yp, Nx, Nxp, mode == 1 ? NULL : &m);
}

int ref_gaussj(double* a, int n, double* b, int m) {
  double** A = dmatrix(1, n, 1, n);
  double** B = dmatrix(1, n, 1, m);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) A[i + 1][j + 1] = a[i * n + j];
    for (int j = 0; j < m; j++) B[i + 1][j + 1] = b[i * m + j];
  }
  int rc = gaussj(A, n, B, m);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) a[i * n + j] = A[i + 1][j + 1];
    for (int j = 0; j < m; j++) b[i * m + j] = B[i + 1][j + 1];
  }
  free_dmatrix(A, 1, n, 1, n);
  free_dmatrix(B, 1, n, 1, m);
  return rc;
}

int ref_adm_chen(void (*f)(int, double*, double*), double* x, double tol, int maxIteration, int n,
                 double lmd, int nn, int Final) {
  return adm_chen(f, x, tol, maxIteration, n, lmd, nn, Final != 0);
}

// f is 0-based (n, in[0..n-1], out[0..n-1]); broydn/fdjac call with 1-based arrays.
static void (*g_f0)(int, double*, double*);
static void tramp1(int n, double* in, double* out) { g_f0(n, in + 1, out + 1); }

int ref_broydn(void (*f)(int, double*, double*), double* x, int n, double tolf, double* err_out, int* jc_io) {
  g_f0 = f;
  qt = dmatrix(1, n, 1, n);
  r = dmatrix(1, n, 1, n);
  d = dvector(1, n);
  jc = jc_io ? *jc_io : 0;
  err = tolf;
  funcerr = 0;
  int check = 1;
  broydn(x - 1, n, &check, tramp1);
  if (err_out) *err_out = err;
  if (jc_io) *jc_io = jc;
  free_dmatrix(qt, 1, n, 1, n);
  free_dmatrix(r, 1, n, 1, n);
  free_dvector(d, 1, n);
  return check;
}

// The reference's broydn handed a RAW residual callback of the reference's own shape
// void f(int n, double in[1..n], double out[1..n]) (broydn.c:44-46, fdjac.c:10-11) — e.g. the address of
// scftb_callback_nr1 — with no adapter in between: this is the call drivescft.cc:301 / 1D_FEM.c:356 makes.
// x is 0-based here and shifted the way NR callers do (x-1).
int ref_broydn_raw(void (*vecfunc)(int, double[], double[]), double* x, int n, double tolf, double* err_out, int* jc_io) {
  qt = dmatrix(1, n, 1, n);
  r = dmatrix(1, n, 1, n);
  d = dvector(1, n);
  jc = jc_io ? *jc_io : 0;
  err = tolf;
  funcerr = 0;
  int check = 1;
  broydn(x - 1, n, &check, vecfunc);
  if (err_out) *err_out = err;
  if (jc_io) *jc_io = jc;
  free_dmatrix(qt, 1, n, 1, n);
  free_dmatrix(r, 1, n, 1, n);
  free_dvector(d, 1, n);
  return check;
}
}
