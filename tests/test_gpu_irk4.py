"""IRK4 (2-stage Gauss-Legendre, the reference's production scheme, scft.cc:671-693) on the GPU:
the only scheme a reference artefact pins.  The GPU solves one complex tridiagonal system per step;
the oracle solves the reference's real 2n x 2n block system with a band LU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-10


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def test_reference_fixture_residual_and_free_energy(sb, oracle, fixtures):
    """DEALII_SCFT/inputFiles/N=33_for_read.txt: converged solution of the reference's own run
    (header ERROR= 1.422819e-09, mean_field_free_energy 0.001945037280931)."""
    eng = sb.Engine(33, nsteps=2048, scheme=sb.IRK4_CONSISTENT)
    out = eng.residual(fixtures["n33_eta"][1:-1])
    assert np.abs(out).max() < 2e-9
    assert eng.free_energy() == pytest.approx(float(fixtures["n33_F"]), abs=1e-15)
    # wall values are the natural-spline extrapolation the reference wrote into the file
    ef = eng.eta_full()
    assert abs(ef[0] - fixtures["n33_eta"][0]) < 1e-14 and abs(ef[-1] - fixtures["n33_eta"][-1]) < 1e-14
    eng.close()


@pytest.mark.parametrize("N,nsteps", [(33, 2048), (17, 16), (65, 256), (129, 64), (257, 64), (300, 32), (513, 64),
                                      (1025, 2048), (1500, 32), (2049, 16)])
def test_irk4_matches_oracle_block_lu(sb, oracle, fixtures, N, nsteps):
    rng = np.random.default_rng(N + nsteps)
    x = oracle.mesh_uniform(N)
    f0 = oracle.f0_given(x)
    eng = sb.Engine(N, nsteps=nsteps, scheme=sb.IRK4_CONSISTENT)
    cases = [rng.standard_normal(N - 2) * 3.0, np.zeros(N - 2)]
    if N == 33:
        cases += [fixtures["n33_eta"][1:-1], fixtures["res32_eta"][1:-1]]
    if N == 1025:
        cases = [fixtures["res1024_eta"][1:-1]]
    for em in cases:
        out = eng.residual(em)
        ref = oracle.residual(oracle.eta_full(x, em), f0, scheme=oracle.IRK4_CONSISTENT, nsteps=nsteps)
        scale = np.abs(ref["phi"]).max()
        assert np.abs(eng.phi() - ref["phi"]).max() < REL * scale
        assert abs(eng.Q() - ref["Q"]) < REL * abs(ref["Q"])
        assert np.abs(out - ref["out"]).max() < REL * scale
    eng.close()


def test_irk4_history_and_batch(sb, oracle):
    N, n, B = 65, 64, 9
    rng = np.random.default_rng(2)
    eng = sb.Engine(N, nsteps=n, scheme=sb.IRK4_CONSISTENT, max_batch=B, store_history=True)
    taus = np.linspace(0.45, 0.6, B)
    Ls = np.linspace(3.4, 4.0, B)
    etas = rng.standard_normal((B, N - 2))
    for p in range(B):
        eng.set_problem(p, taus[p], Ls[p])
    eng.residual(etas)
    for p in (0, 4, 8):
        x = oracle.mesh_uniform(N, Ls[p])
        ref = oracle.residual(oracle.eta_full(x, etas[p]), oracle.f0_given(x, taus[p]), scheme=2, nsteps=n, L=Ls[p],
                              want_hist=True)
        assert np.abs(eng.q_history(p) - ref["hist"]).max() < 1e-12
        assert np.abs(eng.phi(p) - ref["phi"]).max() < REL
    eng.close()


def test_irk4_nonuniform_mesh(sb, oracle, fixtures):
    x = fixtures["matlab59_x"].copy()
    x[-1] = max(x[-1], x[-2] + 1e-3)
    em = fixtures["matlab59_eta"][1:-1]
    eng = sb.Engine(59, nsteps=64, scheme=sb.IRK4_CONSISTENT, tau=0.5302, L=x[-1], x=x)
    eng.residual(em)
    ref = oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x, 0.5302), scheme=2, nsteps=64, L=x[-1], x=x)
    assert np.abs(eng.phi() - ref["phi"]).max() < REL * np.abs(ref["phi"]).max()
    eng.close()


def test_broydn_reproduces_reference_convergence_from_fixture(sb, oracle, fixtures, capfd):
    """the deal.II driver flow (drivescft.cc:270-301): read N=33_for_read.txt, broydn with TOLF 1e-14
    on the IRK4 residual.  The file is already converged to 1.4e-9, so Broyden must leave the field
    essentially where it is and report an error norm of that size or smaller."""
    import ctypes as C
    N = 33
    L = sb.lib()
    eng = sb.Engine(N, nsteps=2048, scheme=sb.IRK4_CONSISTENT, max_batch=N - 2)
    eng.bind_global()
    x = fixtures["n33_eta"][1:-1].copy()
    chk, err, jc = C.c_int(1), C.c_double(1e-14), C.c_int(0)
    rc = L.scftb_broydn(L.scftb_callback_c0, x.ctypes.data_as(C.POINTER(C.c_double)), N - 2, C.byref(chk), C.byref(err),
                        C.byref(jc))
    assert rc == 0
    assert err.value < 3e-9
    assert np.abs(x - fixtures["n33_eta"][1:-1]).max() < 1e-5
    out = eng.residual(x)
    eng.residual(x)
    assert eng.free_energy() == pytest.approx(float(fixtures["n33_F"]), abs=2e-12)
    eng.close()


@pytest.mark.parametrize("nsteps,store", [(2048, False), (64, False), (64, True), (33, False)])
def test_irk4_tensor_memory_kernel_sweep(sb, oracle, fixtures, nsteps, store):
    """march_irk4_tm_kernel (complex coefficients in tensor memory; selected for 513..1024 unknowns, uniform mesh,
    max_batch > 148): 300 sweep problems at N = 1025 — more than the 296 resident CTA slots, so slots are reused — with
    per-problem (tau, L); problems around the wave boundaries against the oracle's block-LU march, lean and full history,
    even and odd step counts"""
    from scft_b200 import sweep
    N, P = 1025, 300
    taus, Ls, eta = sweep.make_sweep(0, P, fixtures["res1024_eta"][1:-1])
    quad = sb.QUAD_ROMBERG if nsteps in (2048, 64) else sb.QUAD_TRAPEZOID
    eng = sb.Engine(N, nsteps=nsteps, scheme=sb.IRK4_CONSISTENT, max_batch=P, store_history=store, quadrature=quad)
    assert "march_irk4_tm_kernel" in eng.kernel_name() and eng.slots() == 296
    for p in range(P):
        eng.set_problem(p, taus[p], Ls[p])
    out = eng.residual(eta)
    for p in ((0, 147, 148, 295, 296, 299) if nsteps != 2048 else (0, 296, 299)):
        x = oracle.mesh_uniform(N, Ls[p])
        ref = oracle.residual(oracle.eta_full(x, eta[p]), oracle.f0_given(x, taus[p]), scheme=oracle.IRK4_CONSISTENT, nsteps=nsteps,
                              L=Ls[p], quadrature=oracle.QUAD_ROMBERG if quad == sb.QUAD_ROMBERG else oracle.QUAD_TRAPEZOID,
                              want_hist=store)
        scale = np.abs(ref["phi"]).max()
        assert np.abs(eng.phi(p) - ref["phi"]).max() < REL * scale
        assert abs(eng.Q(p) - ref["Q"]) < REL * abs(ref["Q"])
        assert np.abs(out[p] - ref["out"]).max() < REL * scale
        if store and p in (0, 299):
            assert np.abs(eng.q_history(p) - ref["hist"]).max() < 1e-11
    eng.close()
