/*
 * scft_b200.h — C ABI of the B200-native SCFT propagator engine.
 *
 * Drop-in boundary for the hot path of giantsda/SCFT: the residual evaluation
 *      eta (field on interior nodes)  ->  phi_0 - phi   (density mismatch)
 * i.e. the FEM march of the modified diffusion equation for q(x,s), the density quadrature
 * and the field updates that call it.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Plain C: opaque handle, plain pointers and sizes, int status codes (0 = ok).  The library never
 * calls exit().  This header does not depend on include order, but like every system header it
 * must not be included AFTER the reference's nr.h / nrutil.h, which "#define float double"
 * (nr.h:8) — include it first, or #undef float.
 *
 * Index conventions: n = N-2 is the number of INTERIOR nodes; arrays are 0-based
 * unless the name says nr1 (Numerical-Recipes 1-based, in[1..n]).
 */
#ifndef SCFT_B200_H_
#define SCFT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define SCFTB_OK 0
#define SCFTB_ERR_ARG 1        /* bad argument / unsupported size */
#define SCFTB_ERR_CUDA 2       /* CUDA runtime error (see scftb_last_error) */
#define SCFTB_ERR_NAN 3        /* NaN in a residual (ADM_chen_C.c:61-66 exit(1)s; we return) */
#define SCFTB_ERR_NOCONV 4     /* solver hit its iteration limit (adm_chen returns 1) */
#define SCFTB_ERR_STATE 5      /* call order (e.g. callback without a bound engine) */

/* ---- discretisation of the contour march --------------------------------------------------- */
#define SCFTB_IE_ROWSCALE 0     /* 1D_FEM.c:95-186 / Matlab_files/simple_FEM_1D_transient.m:34-91 */
#define SCFTB_IE_CONSISTENT 1   /* deal.II A,B,C (scft.cc:643-656) with an implicit-Euler step    */
#define SCFTB_IRK4_CONSISTENT 2 /* 2-stage Gauss-Legendre, scft.cc:671-693 + drivescft.cc:130-146 */
#define SCFTB_QUAD_ROMBERG 0    /* romint.c:21-57 (drivescft.cc:192) */
#define SCFTB_QUAD_TRAPEZOID 1  /* simple_FEM_1D_transient.m:120-124 */

typedef struct scftb_engine scftb_engine;

/* Replaces the state the reference keeps in SCFT::HeatEquation<2> (SCFT.h:117-155; ctor
 * scft.cc:22-37: tau, N, total_time_step, L) and the literals of 1D_FEM.c:60-61. */
typedef struct {
  int scheme;        /* SCFTB_IE_* / SCFTB_IRK4_* */
  int N;             /* nodes of the 1D mesh (elements m = N-1) */
  int nsteps;        /* contour steps n; the reference's total_time_step is n+1 */
  int quadrature;    /* SCFTB_QUAD_* */
  double sign;       /* +1: out = phi0 - phi (drivescft.cc:212); -1: phi - phi0 (1D_FEM.c:276) */
  int max_batch;     /* capacity: number of independent problems held by the engine (>= 1) */
  int device;        /* CUDA device ordinal */
  int store_history; /* 1: keep every q(x,s_j) slice of every problem (solution_store,
                        scft.cc:30); 0: keep only the half the fused quadrature re-reads */
} scftb_config;

int scftb_create(const scftb_config *cfg, scftb_engine **out);
int scftb_destroy(scftb_engine *e);
const char *scftb_last_error(void);
/* number of kernels this library launched in the PROCESS (all engines, all host threads) since the last reset
 * (bench evidence) */
long scftb_launch_count(int reset);

/* Physical parameters of problem p (all problems when p < 0): surface-layer width tau, film
 * thickness L and, optionally, N node coordinates (NULL = uniform mesh, drivescft.cc:91-98).
 * Computes phi_0 = f0_given on the nodes (scft.cc:188-215; 1D_FEM.c:305-319). */
int scftb_set_problem(scftb_engine *e, int p, double tau, double L, const double *x);

/* ---- the hot path -------------------------------------------------------------------------- */
/* One residual evaluation of problem 0, host buffers.  Replaces
 * SCFT::HeatEquation<2>::run(double*) (drivescft.cc:81-216) and
 * simple_FEM_1D_transient(int,double*,double*) (1D_FEM.c:47-287), 0-based. */
int scftb_residual(scftb_engine *e, const double *eta_mid, double *out);
/* nprob independent evaluations (parameter sweep, or the n columns of fdjac.c:18-34), host buffers
 * eta_mid[nprob][N-2], out[nprob][N-2]; H2D and D2H copies are part of the call. */
int scftb_residual_batch(scftb_engine *e, int nprob, const double *eta_mid, double *out);
/* Same with DEVICE buffers on a caller stream (cudaStream_t passed as void*); asynchronous. */
int scftb_residual_batch_device(scftb_engine *e, int nprob, const double *d_eta_mid, double *d_out,
                                void *stream);

/* Results of the last evaluation of problem p (host buffers). */
int scftb_get_phi(scftb_engine *e, int p, double *phi /* N */);         /* f0, drivescft.cc:185-193 */
int scftb_get_Q(scftb_engine *e, int p, double *Q);                      /* (1/L) int q(x,1) dx */
int scftb_get_f0_given(scftb_engine *e, int p, double *f0 /* N */);     /* scft.cc:188-215 */
int scftb_get_eta_full(scftb_engine *e, int p, double *eta /* N */);    /* scft.cc:452-490 */
/* q history, row-major [N][nsteps+1] like solution_store rows 1..N (scft.cc:30,
 * drivescft.cc:123-152).  Needs store_history = 1. */
int scftb_get_q_history(scftb_engine *e, int p, double *hist);
/* Mean-field free energy per segment of the last evaluated field of problem p
 * (scft.cc:404-450 via print_and_save_yita_1D scft.cc:271-291); f0bar <= 0 computes it the way
 * testFiBar.cc:19-50 does. */
int scftb_free_energy(scftb_engine *e, int p, double f0bar, double *F);

/* ---- the reference-shaped residual callback  void f(int n, double* in, double* out) ---------
 * scftb_bind_global plays the role of the global `heat_equation_solver` (drivescft.cc:50);
 * scftb_callback_nr1 replaces SCFT_wrapper under #define BROYDN (drivescft.cc:218-243) and
 * simple_FEM_1D_transient as passed to broydn (1D_FEM.c:356): arrays in[1..n], out[1..n];
 * scftb_callback_c0 replaces SCFT_wrapper for adm_chen (0-based).  On a CUDA failure they set
 * scftb_funcerr (the analogue of broydn.c:25 funcerr) instead of aborting.  The binding (and scftb_broydn's retained
 * QR factors, the reference's caller-owned qt/r/d) is per HOST THREAD: each thread binds the engine it drives.
 * March launches of one engine never overlap, whatever streams they are issued on: its history buffers belong to the
 * resident CTA slots, so a launch on another stream is ordered after the previous one (cudaStreamWaitEvent). */
int scftb_bind_global(scftb_engine *e);
void scftb_callback_nr1(int n, double *in, double *out);
void scftb_callback_c0(int n, double *in, double *out);
/* fixed-point image x + (phi0 - phi) for adm (comment at drivescft.cc:212), 0-based */
void scftb_callback_fixedpoint_c0(int n, double *in, double *out);
extern int scftb_funcerr;

/* ---- field updates -------------------------------------------------------------------------- */
typedef void (*scftb_func)(int n, double *in, double *out);

/* Host-flow solvers with the reference argument lists (generic callback, one problem). */
/* adm_chen (NR_chen.h:30-34, ADM_chen_C.c:18-147); returns 0 converged / 1 not / SCFTB_ERR_NAN */
int scftb_adm_chen(scftb_func f, double *x_old, double tol, int maxIteration, int n, double lmd,
                   int nn, int Final);
/* adm (adm.c:24-313) for x = f(x); 0-based (flag 0); *check 0 converged / 1 failed */
int scftb_adm(scftb_func f, double *x, int n, int *check, int maxits);
/* broydn (broydn.c:44-292), x 0-based here; tolf in / achieved max|f| out via *err; jc as in
 * broydn.c:26-27.  When f == scftb_callback_c0 the finite-difference Jacobian (fdjac.c:18-34)
 * is evaluated as ONE batch of n residuals on the device, all with the parameters of problem 0 of the bound engine
 * (the problem the callback evaluates) — safe on a sweep engine whose slots hold different (tau, L). */
int scftb_broydn(scftb_func f, double *x, int n, int *check, double *err, int *jc);

/* broydn with EVERYTHING on the device: finite-difference Jacobian as one batched launch, Householder QR,
 * Q^T, Givens rank-one updates, triangular solves and line-search vectors stay in HBM; the host steers with a
 * few scalars per step.  Solves problem 0 of the engine; the Jacobian columns are evaluated with problem 0's
 * (tau, L, mesh) in batches of max_batch fields, whatever parameters the other slots of the engine hold (their phi, Q,
 * eta_full are not touched).  x[N-2] host in/out; check / err / jc as scftb_broydn. */
int scftb_broydn_device(scftb_engine *e, double *x, int *check, double *err, int *jc);
/* Same with options.  SCFTB_BROYDN_KEEP_TRIAL: when the line search stops on its step-size test
 * (lnsrch.c: alam < alamin) the reference resets x to the previous iterate but reports the residual norm of
 * the trial point, so a "converged" return can carry a field whose own residual is above the tolerance
 * (broydn.c:215-228).  With this flag a trial point that meets the tolerance is returned itself, so that
 * *err is the residual norm of the returned x. */
#define SCFTB_BROYDN_KEEP_TRIAL 1
int scftb_broydn_device_ex(scftb_engine *e, double *x, int *check, double *err, int *jc, int flags);

/* ---- two-species (AB diblock) extension: q and q+ as separate sweeps ------------------------------
 * Not in the reference (its melt is one species and uses q+(x,s) = q(x,1-s), drivescft.cc:189-190); it is the
 * diblock model of WQ-HardSurf.pdf II.B that gives sweeps their chi N axis.  The chain has an A
 * block of fraction fA (fA*nsteps must be an integer number of contour steps) and a B block; q is marched from the
 * A end and q+ from the B end with the engine's implicit-Euler scheme, every slice of q is kept in HBM and the
 * q+ sweep accumulates phi_A = int_0^fA q q+ ds and phi_B = int_fA^1 q q+ ds on the fly (the engine's quadrature
 * on every block that has 2^k >= 16 steps, else the trapezoid rule).  Unknowns and residual of problem p:
 *     w[2][N-2]   = (eta_A, eta_B) on the interior nodes
 *     out[2][N-2] = ( sign*(phi_0 - phi_A - phi_B),  eta_A - eta_B - chiN*(phi_B - phi_A) )
 * (same gauge as the reference residual: no 1/Q).  With eta_A == eta_B the first half is the reference residual.
 * scftb_set_diblock: fA is a property of the engine, chiN of problem p (all problems when p < 0). */
int scftb_set_diblock(scftb_engine *e, int p, double fA, double chiN);
int scftb_residual_ab(scftb_engine *e, const double *w, double *out);
int scftb_residual_ab_batch(scftb_engine *e, int nprob, const double *w, double *out);
int scftb_get_phi_ab(scftb_engine *e, int p, double *phiA, double *phiB);
/* residual callback of the bound engine for the generic solvers, n = 2*(N-2), 0-based */
void scftb_callback_ab_c0(int n, double *in, double *out);

/* Device-resident batched Anderson mixing (adm_chen semantics) on the engine's problems:
 * x[nprob][N-2] host in/out.  All nprob problems iterate in lock-step, each with its own
 * history, Gram matrix, gaussj solve and relaxation; nothing but the per-problem error norms
 * leaves the device between iterations.  iters_out/err_out (may be NULL): per problem. */
int scftb_adm_chen_batch(scftb_engine *e, int nprob, double *x, double tol, int maxIteration,
                         double lmd, int nn, int Final, int *iters_out, double *err_out);

/* adm (adm.c:24-313) for a batch, device-resident: the fixed-point map is x -> x + (phi0 - phi)
 * (scftb_callback_fixedpoint_c0), history ring of NRMAX = 10 (adm.c:6), lambda = 0.05 then 1 - 0.95^its (adm.c:140,151),
 * gaussj with the zero-pivot nudge of the root gaussj.c:38, TOLF = 1e-10 (adm.c:29).  Bit-identical to scftb_adm driven by
 * scftb_callback_fixedpoint_c0.  At most maxits evaluations per problem; 0 when every problem converged. */
int scftb_adm_batch(scftb_engine *e, int nprob, double *x, int maxits, int *iters_out, double *err_out);

/* The same, one iteration at a time, for callers that keep the fields on the device (sweeps,
 * benchmarks).  A mixer owns the per-problem history rings X,Y (ADM_chen_C.c:49-50), lk and the
 * restart index.  scftb_mixer_iterate_device issues Y_k = F(X_k) and the Anderson update on the
 * caller's stream and returns immediately; converged problems are frozen and skipped. */
typedef struct scftb_mixer scftb_mixer;
int scftb_mixer_create(scftb_engine *e, int nprob, double tol, double lmd, int nn, int Final, scftb_mixer **out);
/* adm semantics instead of adm_chen (see scftb_adm_batch); all other mixer calls apply unchanged */
int scftb_adm_mixer_create(scftb_engine *e, int nprob, scftb_mixer **out);
int scftb_mixer_destroy(scftb_mixer *m);
int scftb_mixer_reset(scftb_mixer *m, const double *x, int x_is_device, void *stream);
/* freeze = 1 (default): converged / NaN problems are skipped from then on, as adm_chen returns or
 * exits for them; 0: every problem is evaluated in every iteration (fixed-iteration benchmarks). */
int scftb_mixer_set_freeze(scftb_mixer *m, int freeze);
int scftb_mixer_iterate_device(scftb_mixer *m, void *stream);
/* done[p]: 0 running, 1 converged, 2 NaN; iters[p]: iteration index at which it stopped; err[p]: last max|F| */
int scftb_mixer_status(scftb_mixer *m, void *stream, int *done, int *iters, double *err);
int scftb_mixer_get_x(scftb_mixer *m, void *stream, double *x /* host [nprob][N-2] */);
/* residuals Y_k = F(X_k) of iteration k (ADM_chen_C.c:50,57), still in the ring while k > iterations issued - nn - 2 */
int scftb_mixer_get_y(scftb_mixer *m, void *stream, int k, double *y /* host [nprob][N-2] */);

/* ---- preconditioned Anderson mixing and the continuation sweep solver (pmixer.cu) ------------------------------
 * B200-first replacement of the reference's staged adm_chen schedule (drivescft.cc:294-298) for batches: the adm_chen
 * update (ADM_chen_C.c:86-123) applied to the preconditioned residual -(I + Psi^-1 (1/2)(-Lap_h) Psi^-1)(phi0 - phi),
 * Psi = diag(sqrt(phi)), relaxation 1; convergence is tested on the raw residual max|phi0 - phi| < tol exactly like
 * adm_chen.  All problems iterate in lock-step on the device; done[p]: 0 running, 1 converged, 2 NaN start field. */
typedef struct scftb_pmixer scftb_pmixer;
int scftb_pmixer_create(scftb_engine *e, int nprob, double tol, int nn /* window, <= 16 */, double cap /* <= 0: 2.0 */,
                        scftb_pmixer **out);
int scftb_pmixer_destroy(scftb_pmixer *m);
int scftb_pmixer_reset(scftb_pmixer *m, const double *x, int x_is_device, void *stream);
int scftb_pmixer_iterate_device(scftb_pmixer *m, void *stream);
int scftb_pmixer_status(scftb_pmixer *m, void *stream, int *done, int *iters, double *err);
/* converged problems: their field; running ones: the best iterate so far.  x [nprob][N-2], host or device */
int scftb_pmixer_get_x(scftb_pmixer *m, void *stream, double *x, int x_is_device);
/* adm_chen-shaped convenience call for a batch: x[nprob][N-2] host in/out; returns 0 / SCFTB_ERR_NOCONV / SCFTB_ERR_NAN.
 * A problem that did not converge gets its BEST iterate back, with that iterate's max|phi0-phi| in err_out. */
int scftb_padm_batch(scftb_engine *e, int nprob, double *x, double tol, int maxIteration, int nn, int *iters_out,
                     double *err_out);
/* refine_mesh (scft.cc:132-169) for a batch of uniform meshes on the device: d_eta[nprob][N-2] -> d_eta_new[nprob][2N-3] */
int scftb_refine_uniform_batch_device(int nprob, int N, const double *d_L, const double *d_eta, double *d_eta_new, void *stream);
/* c[N], f0bar with  free energy = (sum_i c_i eta_i / f0bar / L + log f0bar) / -1000  (scft.cc:404-450 with the
 * 2^18+1-point Romberg rule of scft.cc:271-281 folded into nodal weights; f0bar as testFiBar.cc:19-50); x NULL = uniform */
int scftb_free_energy_weights(int N, const double *x, double tau, double L, double *c, double *f0bar);

/* The reference's driver flow (drivescft.cc:259-322: solve, cut every cell in x, spline transfer, solve again) for a
 * batch of independent problems, N0 -> 2N0-1 -> ... (levels meshes), preconditioned mixing on every level, everything
 * between the start fields and the converged fields in HBM. */
typedef struct scftb_sweep scftb_sweep;
typedef struct {
  int scheme;       /* SCFTB_IE_* / SCFTB_IRK4_* */
  int N0, levels;   /* coarsest mesh and number of meshes: target N = (N0-1) 2^(levels-1) + 1 */
  int nsteps, quadrature;
  double tol;       /* on max|phi0 - phi|, every level (<= 0: 1e-9) */
  int nn;           /* mixing window (<= 0: 10) */
  int maxit;        /* evaluations per level (<= 0: 200) */
  double cap;       /* step cap while max|F| > 1e-2 (<= 0: 2.0) */
  int device;
} scftb_sweep_config;
#define SCFTB_SWEEP_COLS 7
int scftb_sweep_create(const scftb_sweep_config *cfg, int max_prob, scftb_sweep **out);
int scftb_sweep_destroy(scftb_sweep *s);
int scftb_sweep_target_N(scftb_sweep *s);
/* tau[nprob], L[nprob], eta0[nprob][N0-2] (host) -> eta_out[nprob][N_target-2] (host, may be NULL) and
 * rows[nprob][SCFTB_SWEEP_COLS] = { status (0 converged on every level / 1 not / 2 NaN), max|phi0-phi| on the last level
 * reached, evaluations summed over the levels, Q, free energy (f0bar of the problem's own (tau, L)), evaluations on the
 * last level, N of the last level reached }; level_seconds[levels + 1] (may be NULL): wall seconds per level, then the
 * wait for the host threads that prepare the free-energy weights */
int scftb_sweep_solve(scftb_sweep *s, int nprob, const double *tau, const double *L, const double *eta0, double *eta_out,
                      double *rows, double *level_seconds);

/* ---- around the hot path: spline, refinement, result files ------------------- */
/* spline_chen (spline_chen.c:12-106): mode 0 natural (m = 0), 1 not-a-knot (m == NULL), 2 y'' = bc at both
 * ends; evaluates (and extrapolates) at xp.  Solved as a tridiagonal system, not dense gaussj. */
int scftb_spline(const double *x, const double *y, const double *xp, double *yp, int Nx, int Nxp, int mode, double bc);
/* refine_mesh (scft.cc:132-169): every cell cut in x, N -> 2N-1 nodes; the interior field is carried over by a
 * not-a-knot spline through the old interior nodes.  x_new[2N-1], eta_mid_new[2N-3]. */
int scftb_refine_mesh(int N, const double *x, const double *eta_mid, double *x_new, double *eta_mid_new);
/* refine_mesh.m:6-30 (MATLAB prototype): cut the cells whose |d eta/dx| is >= factor x the median (the prototype
 * uses 10) and the two wall cells; not-a-knot transfer.  *N_new <= 2N-1; buffers sized for 2N-1 / 2N-3. */
int scftb_refine_mesh_adaptive(int N, const double *x, const double *eta_mid, double factor, int *N_new, double *x_new,
                               double *eta_mid_new);
/* solution_yita_1D_N=<N>.txt writer (scft.cc:319-337) and reader (read_yita_middle_1D, scft_util.cc:13-41) */
int scftb_write_solution(const char *path, int N, double err, double F, const double *x, const double *eta_full);
int scftb_read_solution(const char *path, int *N, double *x, double *eta, int capacity);
/* detailedsolution_yita_1D_N=<N>.txt (scft.cc:269-312): eta_h sampled on nplot equidistant points (<= 1: 2^18 + 1) */
int scftb_write_detailed_solution(const char *path, int N, double err, double F, const double *x, const double *eta_full,
                                  int nplot);
/* Exp_m*_n2048_IE.res reader (1D_FEM.c:322-342): rows of x/l, phi, eta after 9 header lines */
int scftb_read_res(const char *path, int rows, double *xl, double *phi, double *eta);

/* ---- the 2-D path: Q1 mesh, CSR matrices, Jacobi-PCG per contour step (BASELINE.json configs[3],[4]) ------
 * Structured nx x ny cells on [0,L] x [0,Ly], DOF d = ix*(ny+1)+iy, Dirichlet q = 0 on x = 0 and x = L
 * (scft.cc:599-606), matrices A, B, C of scft.cc:643-656, implicit-Euler step (A + ds(B+C)) q+ = A q solved by
 * conjugate gradients (the role of solve_time_step, scft.cc:698-705).  world > 1: slab partition in x over
 * `world` ranks (one process per GPU), halo exchange and dot products over NCCL. */
typedef struct scftb2d_engine scftb2d_engine;
typedef struct {
  int nx, ny;          /* cells in x and y */
  double L, Ly, tau;   /* domain and surface-layer width (phi0 depends on x only) */
  int nsteps;          /* contour steps */
  int quadrature;      /* SCFTB_QUAD_* */
  double sign;         /* +1: phi0 - phi */
  double rtol;         /* CG stops at ||r|| <= rtol ||b|| (<= 0: 1e-12) */
  int maxit;           /* CG iteration limit per step (<= 0: 100000) */
  int device;          /* CUDA device ordinal of this rank */
  int rank, world;     /* slab partition */
  int store_history;   /* keep all slices instead of the half the quadrature re-reads */
} scftb2d_config;
/* rank 0 obtains the NCCL id and hands the 128 bytes to every rank (e.g. through torch.distributed) */
int scftb2d_nccl_unique_id(char *id128);
int scftb2d_create(const scftb2d_config *cfg, const char *nccl_id128 /* NULL when world == 1 */, scftb2d_engine **out);
int scftb2d_destroy(scftb2d_engine *e);
/* global row range [row0, row0+nrows) owned by this rank */
int scftb2d_rows(scftb2d_engine *e, int *row0, int *nrows);
/* eta on ALL (nx+1)(ny+1) nodes (host) -> sign*(phi0 - phi) on this rank's rows (host, nrows values) */
int scftb2d_residual(scftb2d_engine *e, const double *eta, double *out);
int scftb2d_get_phi(scftb2d_engine *e, double *phi /* nrows */);
int scftb2d_get_stats(scftb2d_engine *e, long long *cg_iterations, double *march_ms);
/* Peer-memory version of the exchange (one box, NVLink): every rank exports the cudaIpc handle (64 bytes) of its
 * exchange buffer, the handles of all ranks are gathered (e.g. torch.distributed.all_gather) and attached; from then
 * on scftb2d_residual runs the whole march as ONE persistent kernel per rank that stores halo columns straight into
 * the neighbours' vectors and all-reduces the dot products through peer-mapped slots. */
int scftb2d_p2p_handle(scftb2d_engine *e, char *handle64);
int scftb2d_p2p_attach(scftb2d_engine *e, const char *handles /* world x 64 bytes */);
/* unmap the peers; call on every rank and synchronise the ranks BEFORE scftb2d_destroy frees the exported memory */
int scftb2d_p2p_detach(scftb2d_engine *e);
/* CSR view (global column indices) of this rank's rows of T = A + ds(B+C) and A after the last assembly:
 * rowptr[nrows+1], colind/valT/valA[<= 9*nrows] */
int scftb2d_export_csr(scftb2d_engine *e, int *rowptr, int *colind, double *valT, double *valA);

/* ---- measurement hooks ---------------------------------------------------------------------- */
/* When on, every march-kernel launch is bracketed by CUDA events on its launching stream;
 * scftb_get_march_ms returns the summed device time and the number of launches since the last call. */
int scftb_set_timing(scftb_engine *e, int on);
int scftb_get_march_ms(scftb_engine *e, double *total_ms, int *count);
/* resident CTA slots of the march kernel on this device (= problems per wave; lean history is kept per slot) */
int scftb_get_slots(scftb_engine *e, int *slots);
/* name and CTA shape of the march kernel this engine launches (bench evidence) */
int scftb_get_kernel_name(scftb_engine *e, char *buf, int len);

#ifdef __cplusplus
}
#endif
#endif /* SCFT_B200_H_ */
