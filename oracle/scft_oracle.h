/*
 * scft_oracle.h — CPU oracle for the SCFT propagator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (scft_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * It is a plain-C restatement of the algorithm of giantsda/SCFT; every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 * All arrays are 0-based and all arithmetic is fp64 (the reference #defines
 * float to double, nr.h:8).
 *
 * Parity status:
 *   IRK4_CONSISTENT + Romberg + free energy : pinned by the reference fixture
 *       DEALII_SCFT/inputFiles/N=33_for_read.txt (residual <= 2e-9, F to 1e-15)
 *   romint / spline / gaussj / adm / adm_chen / broydn : pinned against the
 *       reference's own C compiled into oracle/_ref (tests/test_oracle_vs_ref.py)
 *   IE_ROWSCALE : pinned by Matlab_files/inputFiles/solution_matlab_N=33, a converged solution of
 *       Matlab_files/simple_FEM_1D_transient.m (2049 steps of dt = 1/2048, trapezoid): residual 2.3e-7 at the
 *       MATLAB run's tolerance 1e-7 (tests/test_oracle_golden.py; 1D_FEM.c itself needs PETSc and records no output)
 *   IE_CONSISTENT, Q, two-species : PARITY UNPINNED by any reference artefact (oracle only)
 */
#ifndef SCFT_ORACLE_H_
#define SCFT_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_IE_ROWSCALE = 0,     /* 1D_FEM.c:95-186, simple_FEM_1D_transient.m:34-91 */
  ORC_IE_CONSISTENT = 1,   /* deal.II A,B,C (scft.cc:643-656) with an IE step   */
  ORC_IRK4_CONSISTENT = 2  /* scft.cc:671-693 + drivescft.cc:130-146            */
};
enum { ORC_QUAD_ROMBERG = 0, ORC_QUAD_TRAPEZOID = 1 };
enum { ORC_SPLINE_NATURAL = 0, ORC_SPLINE_NOTAKNOT = 1, ORC_SPLINE_GIVEN = 2 };

typedef struct {
  int scheme;       /* ORC_IE_* / ORC_IRK4_* */
  int N;            /* number of nodes (elements m = N-1) */
  int nsteps;       /* contour steps n (history has n+1 slices) */
  int quadrature;   /* ORC_QUAD_* */
  double sign;      /* +1: out = f0_given - f0 (drivescft.cc:212); -1: f0 - f0_given (1D_FEM.c:276) */
  double L;         /* film thickness; used when x == NULL (uniform mesh) */
  const double *x;  /* N node coordinates or NULL for uniform h = L/(N-1) */
} orc_config;

/* Romberg integration of m+1 samples, m = 2^k >= 16 (romint.c:21-57, polint.c:5-42). */
double orc_romint(const double *f, int m, double hh);
/* The linear functional orc_romint(.,m,hh) as an explicit weight vector w[0..m]. */
void orc_romberg_weights(int m, double hh, double *w);

/* phi_0(x) target profile (scft.cc:188-215; 1D_FEM.c:305-319). */
void orc_f0_given(int N, const double *x, double tau, double *f0);
/* mean of phi_0 by Romberg on 2^16+1 points (testFiBar.cc:19-50). */
double orc_f0bar(double tau, double L);

/* Cubic spline (spline_chen.c:12-106) solved as a banded system instead of dense gaussj.
 * mode NATURAL uses second derivative *m==0 semantics (m=0), GIVEN uses bc as y''. */
void orc_spline(const double *x, const double *y, const double *xp, double *yp,
                int Nx, int Nxp, int mode, double bc);

/* eta on all N nodes from the N-2 interior values: natural spline through the interior
 * knots evaluated (extrapolated) at the two wall nodes (scft.cc:452-490). */
void orc_eta_full(int N, const double *x, const double *eta_mid, double *eta_full);

/* One residual evaluation (the hot path).  eta_full: N values.  Outputs (any may be NULL):
 *   out_mid[N-2]  sign*(f0_given - phi) on interior nodes
 *   phi[N]        density
 *   q_hist[N*(nsteps+1)]  row-major, row i = node i, column j = contour step j
 *   Q             (1/L) * int q(x,1) dx  (trapezoid on the mesh; not in the reference) */
int orc_residual(const orc_config *cfg, const double *eta_full, const double *f0_given,
                 double *out_mid, double *phi, double *q_hist, double *Q);

/* Two-species (AB diblock) extension of the residual — not in the reference, parity by oracle only.
 * etaA/etaB: N values each (full fields); jf: contour steps of the A block (0 < jf < nsteps);
 * out[2*(N-2)]: sign*(phi0 - phiA - phiB) then etaA - etaB - chiN*(phiB - phiA); IE schemes only. */
int orc_residual_ab(const orc_config *cfg, const double *etaA, const double *etaB, int jf, double chiN,
                    const double *f0_given, double *out, double *phiA, double *phiB, double *Q);

/* Mean-field free energy per segment (scft.cc:404-450 via :271-291). */
double orc_free_energy(int N, const double *x, const double *eta_full, double tau, double L,
                       double f0bar, int nplot);

/* Gauss-Jordan with full pivoting, a is n*n row-major, b is n*m row-major.
 * variant 0: DEALII_SCFT/src/gaussj.c:7-78 (returns 1 on a zero pivot);
 * variant 1: root gaussj.c:7-60 (nudges a zero pivot by 1e-18, returns 0). */
int orc_gaussj(double *a, int n, double *b, int m, int variant);

typedef void (*orc_func)(int n, double *in, double *out);
/* Anderson mixing for F(x)=0 (ADM_chen_C.c:18-147). trace (optional, maxIteration+2 entries)
 * receives the max-norm error of every iteration; *iters the iteration count. Returns 0/1. */
int orc_adm_chen(orc_func f, double *x, double tol, int maxIteration, int n, double lmd, int nn,
                 int Final, double *trace, int *iters);
/* Anderson mixing for x=f(x) (adm.c:24-313), 0-based, err=1e-10, NRMAX=10. */
int orc_adm(orc_func f, double *x, int n, int maxits, double *trace, int *iters);

/* scft_fast.c — the "fair CPU" baseline: Thomas factorisation once per field, half history, VL problems interleaved
 * for SIMD, POSIX threads over groups of problems.  IE schemes, uniform meshes.  Checked against orc_residual in tests/. */
int orc_fast_sweep(int nprob, int N, int nsteps, int scheme, int quadrature, double sign, const double *tau,
                   const double *L, const double *eta_mid, double *out, double *phi_out, double *Q_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
