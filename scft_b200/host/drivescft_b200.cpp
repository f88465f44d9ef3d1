// drivescft_b200.cpp — host driver with the control flow of the reference's two mains, on top of the
// C ABI (include/scft_b200.h):
//   --flow dealii  DEALII_SCFT/drivescft.cc:259-322: read solution file -> [solve -> save -> refine] x levels
//   --flow 1dfem   1D_FEM.c:289-370: N=33, eta0 from Exp_m32_n2048_IE.res, row-scaled IE, broydn(err=1e-8)
// Solvers: broydn (#define BROYDN, drivescft.cc:45,301) or the staged adm_chen schedule (drivescft.cc:294-298).
// Writes solution_yita_1D_N=<N>.txt in the reference format after every level (scft.cc:319-337).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scft_b200.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CHECK(call)                                                                       \
  do {                                                                                    \
    int _rc = (call);                                                                     \
    if (_rc != SCFTB_OK && _rc != SCFTB_ERR_NOCONV) {                                     \
      fprintf(stderr, "%s failed (%d): %s\n", #call, _rc, scftb_last_error());            \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

int main(int argc, char **argv) {
  std::string flow = "dealii", input, solver = "broydn", scheme_s = "irk4", outdir = ".";
  int levels = 1, nsteps = 2048, device = 0;
  bool recheck = false, detailed = false;
  double tol = -1, tau = 5.30252230020752e-01, L = 3.72374357332160;  // drivescft.cc:269
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "--flow") flow = next();
    else if (a == "--solver") solver = next();
    else if (a == "--scheme") scheme_s = next();
    else if (a == "--levels") levels = atoi(next().c_str());
    else if (a == "--nsteps") nsteps = atoi(next().c_str());
    else if (a == "--tol") tol = atof(next().c_str());
    else if (a == "--tau") tau = atof(next().c_str());
    else if (a == "--L") L = atof(next().c_str());
    else if (a == "--outdir") outdir = next();
    else if (a == "--device") device = atoi(next().c_str());
    else if (a == "--recheck") recheck = true;     // redoFxandshowError.cc:272-290: evaluate a saved solution, report max|F|, rewrite the files
    else if (a == "--detailed") detailed = true;   // also write detailedsolution_yita_1D_N=<N>.txt (scft.cc:293-312), 2^18+1 rows
    else input = a;
  }
  if (input.empty()) {
    fprintf(stderr, "usage: drivescft_b200 <N=33_for_read.txt | Exp_m32_n2048_IE.res> [--flow dealii|1dfem] "
                    "[--scheme irk4|ie|ie_rowscale] [--solver broydn|broydn_dev|adm_chen|adm|padm] [--levels K] [--nsteps n] [--tol t] "
                    "[--recheck] [--detailed]\n");
    return 2;
  }
  int scheme = scheme_s == "irk4" ? SCFTB_IRK4_CONSISTENT : (scheme_s == "ie" ? SCFTB_IE_CONSISTENT : SCFTB_IE_ROWSCALE);
  double sign = +1.0;
  int N = 0;
  std::vector<double> x, eta;  // full mesh / field incl. wall nodes
  if (flow == "1dfem") {       // 1D_FEM.c:292-342
    N = 33; tau = 0.5302; L = 3.72374; scheme = SCFTB_IE_ROWSCALE; sign = -1.0; solver = "broydn";
    if (tol < 0) tol = 1e-8;   // 1D_FEM.c:350
    x.resize(N); eta.assign(N, 0.0);
    std::vector<double> col(N);
    CHECK(scftb_read_res(input.c_str(), N, nullptr, nullptr, col.data()));
    for (int i = 0; i < N; i++) { x[i] = i * L / (N - 1); eta[i] = col[i]; }
  } else {                     // drivescft.cc:270
    CHECK(scftb_read_solution(input.c_str(), &N, nullptr, nullptr, 0));
    x.resize(N); eta.resize(N);
    CHECK(scftb_read_solution(input.c_str(), &N, x.data(), eta.data(), N));
    if (tol < 0) tol = 1e-14;  // drivescft.cc:288
  }
  printf("flow=%s scheme=%s solver=%s N=%d nsteps=%d tol=%g levels=%d\n", flow.c_str(), scheme_s.c_str(), solver.c_str(), N,
         nsteps, tol, levels);

  for (int level = 0; level < levels; level++) {  // drivescft.cc:291-322
    const int n = N - 2;
    scftb_engine *e = nullptr;
    scftb_config cfg = {scheme, N, nsteps, SCFTB_QUAD_ROMBERG, sign, n, device, 0};
    CHECK(scftb_create(&cfg, &e));
    CHECK(scftb_set_problem(e, -1, tau, L, nullptr));
    CHECK(scftb_bind_global(e));
    std::vector<double> xm(eta.begin() + 1, eta.end() - 1), res(n);
    double t0 = now();
    int check = 1, rc = 0;
    if (recheck) {               // no solve: the saved field is evaluated as it is
      check = 0;
    } else if (solver == "padm") {      // preconditioned Anderson mixing (pmixer.cu)
      rc = scftb_padm_batch(e, 1, xm.data(), tol, 400, 10, nullptr, nullptr);
      check = rc == SCFTB_OK ? 0 : 1;
    } else if (solver == "adm") {       // adm.c semantics on the device (fixed TOLF = 1e-10)
      rc = scftb_adm_batch(e, 1, xm.data(), 100000, nullptr, nullptr);
      check = rc == SCFTB_OK ? 0 : 1;
    } else if (solver == "broydn") {
      double err = tol;
      int jc = 0;
      rc = scftb_broydn(scftb_callback_c0, xm.data(), n, &check, &err, &jc);
    } else if (solver == "broydn_dev") {   // the same iteration with Jacobian, QR and updates resident on the device
      double err = tol;
      int jc = 0;
      rc = scftb_broydn_device_ex(e, xm.data(), &check, &err, &jc, SCFTB_BROYDN_KEEP_TRIAL);
    } else {  // staged schedule of drivescft.cc:294-298
      const double st[5][4] = {{1e-1, 200, 0.99, 2}, {1e-3, 300, 0.9, 3}, {1e-7, 800, 0.9, 15}, {1e-7, 1000, 0.9, 30},
                               {1e-7, 10000, 0.1, 50}};
      for (int s = 0; s < 5; s++) {
        double t = st[s][0] < tol ? tol : st[s][0];
        rc = scftb_adm_chen_batch(e, 1, xm.data(), (s == 4 && tol < 1e-7) ? tol : t, (int)st[s][1], st[s][2], (int)st[s][3],
                                  s == 4, nullptr, nullptr);
        if (rc != SCFTB_OK && rc != SCFTB_ERR_NOCONV) break;
      }
      check = rc == SCFTB_OK ? 0 : 1;
    }
    if (rc != SCFTB_OK && rc != SCFTB_ERR_NOCONV && rc != SCFTB_ERR_NAN) { fprintf(stderr, "solver failed: %s\n", scftb_last_error()); return 1; }
    double t_solve = now() - t0;
    // print_and_save_yita_1D (scft.cc:246-339): residual of the final field, free energy, result file
    CHECK(scftb_residual(e, xm.data(), res.data()));
    double emax = 0;
    for (double v : res) emax = std::fmax(emax, std::fabs(v));
    std::vector<double> full(N);
    CHECK(scftb_get_eta_full(e, 0, full.data()));
    double F = 0, Q = 0;
    CHECK(scftb_free_energy(e, 0, 0.0, &F));
    CHECK(scftb_get_Q(e, 0, &Q));
    char path[512];
    snprintf(path, sizeof path, "%s/solution_yita_1D_N=%03d.txt", outdir.c_str(), N);
    CHECK(scftb_write_solution(path, N, emax, F, x.data(), full.data()));
    if (detailed) {
      char dpath[512];
      snprintf(dpath, sizeof dpath, "%s/detailedsolution_yita_1D_N=%03d.txt", outdir.c_str(), N);
      CHECK(scftb_write_detailed_solution(dpath, N, emax, F, x.data(), full.data(), 0));
    }
    printf("level %d: N=%d check=%d Error= %e mean_field_free_energy=%2.15f Q=%2.15f solve_time=%.3fs -> %s\n", level, N, check,
           emax, F, Q, t_solve, path);
    fflush(stdout);
    scftb_destroy(e);
    if (level + 1 < levels) {  // refine_mesh (scft.cc:132-169)
      std::vector<double> xn(2 * N - 1), en(2 * N - 3);
      CHECK(scftb_refine_mesh(N, x.data(), xm.data(), xn.data(), en.data()));
      N = 2 * N - 1;
      x = xn;
      eta.assign(N, 0.0);
      std::copy(en.begin(), en.end(), eta.begin() + 1);
    }
  }
  return 0;
}
