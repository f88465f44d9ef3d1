"""Pin the CPU oracle (oracle/scft_oracle.c) against every golden vector the reference holds for
the hot path and against outputs of the unmodified reference C (ref_outputs.npz)."""
import numpy as np
import pytest

from toys import toyF, toy3, toy2, fp4


def test_romint_matches_reference(oracle, refout):
    f = refout["romint_in"]
    assert oracle.romint(f, 1. / 2048) == pytest.approx(float(refout["romint_out"]), rel=0, abs=1e-16)
    g = np.random.default_rng(1234)
    g.standard_normal(2049)
    big = g.standard_normal(65537)
    assert oracle.romint(big, oracle.L_REF / 65536) == pytest.approx(float(refout["romint65537_out"]), rel=1e-15)


@pytest.mark.parametrize("m", [16, 64, 2048])
def test_romberg_weights_reproduce_romint(oracle, m):
    rng = np.random.default_rng(m)
    w = oracle.romberg_weights(m, 1. / m)
    assert abs(w.sum() - 1.0) < 1e-14
    assert np.allclose(w, w[::-1], rtol=0, atol=1e-18)  # symmetric => the half-history quadrature is exact
    for _ in range(3):
        f = rng.standard_normal(m + 1)
        assert abs(w @ f - oracle.romint(f, 1. / m)) < 5e-16 * np.abs(f).sum() / m + 1e-16


def test_f0bar_matches_hardcoded_constant(oracle):
    # scft.cc:368,413 hard-code the value testFiBar.cc:50 prints
    assert oracle.f0bar() == pytest.approx(0.892581217773656, abs=5e-16)


def test_f0_given_matches_res_phi_column(oracle, fixtures):
    for m in (32, 1024):
        x = fixtures[f"res{m}_xl"] * oracle.L_REF
        assert len(x) == m + 1
        f0 = oracle.f0_given(oracle.mesh_uniform(m + 1))
        assert np.abs(f0 - fixtures[f"res{m}_phi"]).max() < 1e-10
        assert float(fixtures[f"res{m}_mphi"]) == pytest.approx(oracle.f0bar(), rel=2e-11)


def test_free_energy_of_res_eta_matches_testFiBar(oracle, fixtures):
    # testFiBar.cc:107-111 prints integration=0.002016735466532 (.res f=2.01673546641703e-03)
    x = fixtures["res32_xl"] * oracle.L_REF
    g = fixtures["res32_eta"] * oracle.f0_given(x)
    fb = oracle.f0bar()
    F = (oracle.romint(g, oracle.L_REF / 32) / fb / oracle.L_REF + np.log(fb)) / (-1000.)
    assert F == pytest.approx(0.002016735466532, abs=5e-16)
    assert F == pytest.approx(float(fixtures["res32_f"]), rel=1e-10)


def test_irk4_residual_of_converged_fixture(oracle, fixtures):
    """DEALII_SCFT/inputFiles/N=33_for_read.txt is a converged solution of exactly the
    IRK4 + consistent-C + Romberg scheme (header ERROR= 1.422819e-09)."""
    x = oracle.mesh_uniform(33)
    assert np.abs(x - fixtures["n33_x"]).max() < 1e-14
    eta_mid = fixtures["n33_eta"][1:-1]
    ef = oracle.eta_full(x, eta_mid)
    # the file's end values are the natural-spline extrapolation the reference writes (scft.cc:254-257)
    assert abs(ef[0] - fixtures["n33_eta"][0]) < 5e-15 and abs(ef[-1] - fixtures["n33_eta"][-1]) < 5e-15
    f0 = oracle.f0_given(x)
    r = oracle.residual(ef, f0, scheme=oracle.IRK4_CONSISTENT)
    assert np.abs(r["out"]).max() < 2e-9
    # the other schemes / quadratures are NOT what produced the file
    assert np.abs(oracle.residual(ef, f0, scheme=oracle.IE_CONSISTENT)["out"]).max() > 5e-5
    assert np.abs(oracle.residual(ef, f0, scheme=oracle.IRK4_CONSISTENT,
                                  quadrature=oracle.QUAD_TRAPEZOID)["out"]).max() > 1e-6
    assert np.abs(oracle.residual(ef, f0, scheme=oracle.IE_ROWSCALE)["out"]).max() > 5e-2


def test_free_energy_of_converged_fixture(oracle, fixtures):
    x = oracle.mesh_uniform(33)
    ef = oracle.eta_full(x, fixtures["n33_eta"][1:-1])
    assert oracle.free_energy(x, ef) == pytest.approx(float(fixtures["n33_F"]), abs=6e-16)


def test_history_layout_and_symmetry(oracle, fixtures):
    x = oracle.mesh_uniform(33)
    ef = oracle.eta_full(x, fixtures["n33_eta"][1:-1])
    r = oracle.residual(ef, oracle.f0_given(x), scheme=oracle.IE_CONSISTENT, nsteps=64, want_hist=True)
    h = r["hist"]
    assert h.shape == (33, 65)
    assert np.all(h[0] == 0) and np.all(h[-1] == 0) and np.all(h[1:-1, 0] == 1)
    w = oracle.romberg_weights(64, 1. / 64)
    assert np.abs((h * h[:, ::-1]) @ w - r["phi"]).max() < 1e-15


def test_spline_matches_reference(oracle, refout):
    for N in (33, 129):
        x = oracle.mesh_uniform(N)
        yp = oracle.spline(x[1:-1], refout[f"spline_nat{N}_y"], x, oracle.SPLINE_NATURAL)
        assert np.abs(yp - refout[f"spline_nat{N}_yp"]).max() < 2e-13
        # on a uniform mesh the natural-spline extrapolation to the walls is linear
        y = refout[f"spline_nat{N}_y"]
        assert abs(yp[0] - (2 * y[0] - y[1])) < 1e-13
    x, xp = oracle.mesh_uniform(33), oracle.mesh_uniform(65)
    yp = oracle.spline(x[1:-1], refout["spline_nak_y"], xp[1:-1], oracle.SPLINE_NOTAKNOT)
    assert np.abs(yp - refout["spline_nak_yp"]).max() < 1e-12
    xn = refout["spline_nonuni_x"]
    yp = oracle.spline(xn[1:-1], refout["spline_nonuni_y"], xn, oracle.SPLINE_NATURAL)
    assert np.abs(yp - refout["spline_nonuni_yp"]).max() < 1e-11


def test_gaussj_matches_reference_bitwise(oracle, refout):
    rc, ainv, x = oracle.gaussj(refout["gaussj_A"], refout["gaussj_B"])
    assert rc == 0
    assert np.array_equal(ainv, refout["gaussj_Ainv"]) and np.array_equal(x, refout["gaussj_X"])
    rc, _, _ = oracle.gaussj(np.zeros((3, 3)), np.ones((3, 1)), variant=0)
    assert rc == 1  # DEALII gaussj.c:46-50 reports singularity
    rc, _, _ = oracle.gaussj(np.zeros((3, 3)), np.ones((3, 1)), variant=1)
    assert rc == 0  # root gaussj.c:38 nudges the pivot instead


def test_adm_chen_matches_reference_bitwise(oracle, refout):
    rc, x, trace, it = oracle.adm_chen(toyF, [1., 2., 3.], 1e-13, 500, 0.9, 3)
    assert rc == 0 and np.array_equal(x, refout["admchen_toyF_x"])
    assert trace[-1] < 1e-13 and np.abs(toyF(x)).max() < 1e-13
    rc, x, trace, it = oracle.adm_chen(toy3, [1., 2., 3.], 1e-12, 2000, 0.99, 30)
    assert rc == 0 and np.array_equal(x, refout["admchen_toy3_x"])


def test_adm_matches_reference_bitwise(oracle, refout):
    rc, x, trace, it = oracle.adm(toy2, [1., 2.])
    assert rc == 0 and it == 3 and np.array_equal(x, refout["adm_toy2_x"])
    assert np.allclose(x, [-4., 6.], atol=1e-9)
    rc, x, trace, it = oracle.adm(fp4, [1., 2., 3., 4.])
    assert rc == 0 and np.array_equal(x, refout["adm_fp4_x"])


def test_strip_mesh_reduces_to_1d(oracle, fixtures):
    """The deal.II mesh is an (N-1)x1 strip of Q1 cells (drivescft.cc:91-98); its y-invariant
    solution must equal the 1D restatement.  Assemble the strip with a numerical 2x2 Gauss rule
    exactly as scft.cc:643-656 does, constrain x=0,L (scft.cc:599-606), and march IRK4 with a
    sparse LU of the 2*n_dof block matrix (scft.cc:671-695)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    N, n = 33, 128
    L = oracle.L_REF
    x = oracle.mesh_uniform(N)
    ef = oracle.eta_full(x, fixtures["n33_eta"][1:-1])
    hy = L / N
    nd = 2 * N  # dof = 2*i + (0 bottom, 1 top)
    gp = np.array([-1, 1]) / np.sqrt(3)
    A = np.zeros((nd, nd)); B = np.zeros((nd, nd)); Cm = np.zeros((nd, nd))
    for c in range(N - 1):
        hx = x[c + 1] - x[c]
        dofs = [2 * c, 2 * (c + 1), 2 * c + 1, 2 * (c + 1) + 1]  # (x0,y0),(x1,y0),(x0,y1),(x1,y1)
        etal = [ef[c], ef[c + 1], ef[c], ef[c + 1]]
        for gx in gp:
            for gy in gp:
                u, v = (gx + 1) / 2, (gy + 1) / 2
                sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
                gr = np.array([[-(1 - v) / hx, -(1 - u) / hy], [(1 - v) / hx, -u / hy],
                               [-v / hx, (1 - u) / hy], [v / hx, u / hy]])
                jxw = hx * hy / 4
                eq = float(sh @ etal)
                for a in range(4):
                    for b in range(4):
                        A[dofs[a], dofs[b]] += sh[a] * sh[b] * jxw
                        B[dofs[a], dofs[b]] += gr[a] @ gr[b] * jxw
                        Cm[dofs[a], dofs[b]] += sh[a] * sh[b] * eq * jxw
    D = B + Cm
    free = np.array([d for d in range(nd) if d // 2 not in (0, N - 1)])
    A, D = A[np.ix_(free, free)], D[np.ix_(free, free)]
    dt = 1. / n
    c01, c10 = (0.25 - np.sqrt(3) / 6) * dt, (0.25 + np.sqrt(3) / 6) * dt
    blk = sp.csc_matrix(np.block([[A + dt / 4 * D, c01 * D], [c10 * D, A + dt / 4 * D]]))
    lu = spl.splu(blk)
    q = np.ones(len(free))
    hist = [q.copy()]
    for _ in range(n):
        t = -D @ q
        k = lu.solve(np.concatenate([t, t]))
        q = q + 0.5 * dt * (k[:len(free)] + k[len(free):])
        hist.append(q.copy())
    hist = np.array(hist).T  # dof x step
    r = oracle.residual(ef, oracle.f0_given(x), scheme=oracle.IRK4_CONSISTENT, nsteps=n, want_hist=True)
    bottom = hist[0::2]
    top = hist[1::2]
    assert np.abs(bottom - top).max() < 1e-12               # y-invariant
    assert np.abs(bottom - r["hist"][1:-1]).max() < 1e-11   # equals the 1D restatement


# ---- the reference artefact that pins the implicit-Euler / row-scaled scheme -------------------------------------
# Matlab_files/inputFiles/solution_matlab_N=33 is a converged solution of Matlab_files/simple_FEM_1D_transient.m
# (drive_SCFT.m:4-18: tau = 0.5302, L = 3.72374, adm_chen to 1e-7).  That function marches time_step = 2049 steps of
# dt = 1/2048 (:17-18,:85-117: the prototype's 2049-vs-2048 defect) and integrates with the trapezoid rule (:120-127).
# The engine/oracle use ds = 1/nsteps; the MATLAB march is reproduced EXACTLY with nsteps = 2049 by the similarity
#   x -> c x, tau -> c tau, eta -> eta * 2049/2048, phi -> phi * 2049/2048,  c = sqrt(2048/2049)
# (A + ds(B + eta A) scales by c when ds/c^2 and ds*eta are kept; phi_0 depends on x/tau only).
MATLAB_TAU, MATLAB_L = 0.5302, 3.72374


def matlab_scaled_problem(fixtures):
    x, eta = fixtures["matlab33_x"], fixtures["matlab33_eta"]
    c = np.sqrt(2048.0 / 2049.0)
    return x, eta, c


def test_ie_rowscale_pinned_by_the_matlab_converged_solution(oracle, fixtures):
    x, eta, c = matlab_scaled_problem(fixtures)
    xs = oracle.mesh_uniform(33, MATLAB_L * c)
    f0 = oracle.f0_given(x, MATLAB_TAU)
    assert np.abs(oracle.f0_given(xs, MATLAB_TAU * c) - f0).max() < 1e-15
    res = oracle.residual(eta * 2049.0 / 2048.0, f0, scheme=oracle.IE_ROWSCALE, nsteps=2049, L=MATLAB_L * c,
                          quadrature=oracle.QUAD_TRAPEZOID)
    out = f0[1:-1] - res["phi"][1:-1] * 2049.0 / 2048.0
    assert np.abs(out).max() < 3e-7           # the file is converged to adm_chen's 1e-7 (measured 2.33e-7)


def test_oracle_equals_a_dense_restatement_of_the_matlab_function(oracle, fixtures):
    """simple_FEM_1D_transient.m:34-128 restated with dense numpy matrices (inv(D) as the .m file does)"""
    x, eta, c = matlab_scaled_problem(fixtures)
    N, ni, dt = len(x), len(x) - 2, 1.0 / 2048
    a1, a2 = np.diff(x)[:-1], np.diff(x)[1:]
    A, B = np.zeros((ni, ni)), np.zeros((ni, ni))
    for i in range(ni):                                                    # :35-63 (interior rows)
        A[i, i], B[i, i] = a1[i] / 3 + a2[i] / 3, 1 / a1[i] + 1 / a2[i]
        if i > 0:
            A[i, i - 1], B[i, i - 1] = a1[i] / 6, -1 / a1[i]
        if i < ni - 1:
            A[i, i + 1], B[i, i + 1] = a2[i] / 6, -1 / a2[i]
    Dinv = np.linalg.inv(A + dt * (B + np.diag(eta[1:-1]) @ A))            # :66-80
    q, H = np.ones(ni), [np.ones(ni)]
    for _ in range(2049):                                                  # :85-117, time_step = 2049
        q = Dinv @ (A @ q)
        H.append(q.copy())
    H = np.array(H)
    v = H * H[::-1]
    phi = dt * (v[0] / 2 + v[1:-1].sum(axis=0) + v[-1] / 2)                # :120-127
    f0 = oracle.f0_given(x, MATLAB_TAU)
    res = oracle.residual(eta * 2049.0 / 2048.0, f0, scheme=oracle.IE_ROWSCALE, nsteps=2049, L=MATLAB_L * c,
                          quadrature=oracle.QUAD_TRAPEZOID)
    assert np.abs(res["phi"][1:-1] * 2049.0 / 2048.0 - phi).max() < 1e-12
    assert np.abs(f0[1:-1] - phi).max() < 3e-7
