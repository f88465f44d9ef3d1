"""GPU parity of the march kernel (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: relative error <= 1e-10 on phi(x) and Q (fp64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-10


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def eta_cases(fixtures, N, rng):
    """fields on the interior nodes of an N-node mesh"""
    if N == 33:
        yield "converged_N33", fixtures["n33_eta"][1:-1]
        yield "res_m32", fixtures["res32_eta"][1:-1]
    if N == 1025:
        yield "res_m1024", fixtures["res1024_eta"][1:-1]
    yield "zero", np.zeros(N - 2)
    yield "random", rng.standard_normal(N - 2) * 3.0


@pytest.mark.parametrize("scheme", [0, 1])
@pytest.mark.parametrize("N,nsteps", [(33, 2048), (33, 64), (17, 16), (65, 256), (129, 128), (257, 64), (513, 64),
                                      (1025, 2048), (100, 32), (1000, 64), (2049, 32), (3000, 16)])
def test_residual_matches_oracle(sb, oracle, fixtures, scheme, N, nsteps):
    rng = np.random.default_rng(N * 7 + nsteps)
    x = oracle.mesh_uniform(N)
    f0 = oracle.f0_given(x)
    eng = sb.Engine(N, nsteps=nsteps, scheme=scheme)
    assert np.array_equal(eng.f0_given(), f0)
    for name, em in eta_cases(fixtures, N, rng):
        out = eng.residual(em)
        ef = oracle.eta_full(x, em)
        ref = oracle.residual(ef, f0, scheme=scheme, nsteps=nsteps)
        assert rel_err(eng.eta_full(), ef) < 1e-13, name
        assert rel_err(eng.phi(), ref["phi"]) < REL, (name, rel_err(eng.phi(), ref["phi"]))
        assert abs(eng.Q() - ref["Q"]) < REL * abs(ref["Q"]), name
        assert np.abs(out - ref["out"]).max() < REL * np.abs(ref["phi"]).max(), name
    eng.close()


def test_trapezoid_and_sign(sb, oracle, fixtures):
    N, n = 33, 100  # trapezoid does not need n = 2^k
    x = oracle.mesh_uniform(N)
    f0 = oracle.f0_given(x)
    em = fixtures["res32_eta"][1:-1]
    eng = sb.Engine(N, nsteps=n, scheme=0, quadrature=sb.QUAD_TRAPEZOID, sign=-1.0)
    out = eng.residual(em)
    ref = oracle.residual(oracle.eta_full(x, em), f0, scheme=0, nsteps=n, quadrature=1, sign=-1.0)
    assert np.abs(out - ref["out"]).max() < REL
    eng.close()


def test_history_matches_oracle(sb, oracle, fixtures):
    N, n = 65, 128
    x = oracle.mesh_uniform(N)
    rng = np.random.default_rng(5)
    em = rng.standard_normal(N - 2)
    eng = sb.Engine(N, nsteps=n, scheme=1, store_history=True)
    eng.residual(em)
    ref = oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x), scheme=1, nsteps=n, want_hist=True)
    h = eng.q_history()
    assert h.shape == (N, n + 1)
    assert np.abs(h - ref["hist"]).max() < 1e-12
    eng.close()


def test_batch_of_distinct_problems(sb, oracle, fixtures):
    """a (tau, L, eta) sweep: every problem of the batch must equal its own oracle evaluation"""
    N, n, B = 129, 64, 37
    rng = np.random.default_rng(11)
    eng = sb.Engine(N, nsteps=n, scheme=1, max_batch=B)
    taus = np.linspace(0.40, 0.66, B)
    Ls = np.linspace(3.2, 4.2, B)
    etas = rng.standard_normal((B, N - 2)) * 2
    for p in range(B):
        eng.set_problem(p, taus[p], Ls[p])
    out = eng.residual(etas)
    for p in range(B):
        x = oracle.mesh_uniform(N, Ls[p])
        f0 = oracle.f0_given(x, taus[p])
        ref = oracle.residual(oracle.eta_full(x, etas[p]), f0, scheme=1, nsteps=n, L=Ls[p])
        assert rel_err(eng.phi(p), ref["phi"]) < REL
        assert np.abs(out[p] - ref["out"]).max() < REL
        assert abs(eng.Q(p) - ref["Q"]) < REL * ref["Q"]
    eng.close()


def test_nonuniform_mesh(sb, oracle, fixtures):
    for n_nodes in (43, 59):
        x = fixtures[f"matlab{n_nodes}_x"].copy()
        x[-1] = max(x[-1], x[-2] + 1e-3)
        L = x[-1]
        em = fixtures[f"matlab{n_nodes}_eta"][1:-1]
        for scheme in (0, 1):
            eng = sb.Engine(n_nodes, nsteps=64, scheme=scheme, tau=0.5302, L=L, x=x)
            out = eng.residual(em)
            f0 = oracle.f0_given(x, 0.5302)
            ef = oracle.eta_full(x, em)
            ref = oracle.residual(ef, f0, scheme=scheme, nsteps=64, L=L, x=x)
            assert rel_err(eng.eta_full(), ef) < 1e-12
            assert rel_err(eng.phi(), ref["phi"]) < REL
            assert abs(eng.Q() - ref["Q"]) < REL * ref["Q"]
            eng.close()


def test_free_energy_matches_fixture(sb, oracle, fixtures):
    eng = sb.Engine(33, nsteps=64, scheme=1)
    eng.residual(fixtures["n33_eta"][1:-1])
    assert eng.free_energy() == pytest.approx(float(fixtures["n33_F"]), abs=1e-15)
    assert eng.free_energy(f0bar=0.0) == pytest.approx(float(fixtures["n33_F"]), abs=1e-14)
    eng.close()


def test_matlab_converged_solution_is_a_fixed_point_on_the_gpu(sb, oracle, fixtures):
    """Reference artefact for the implicit-Euler / row-scaled scheme: Matlab_files/inputFiles/solution_matlab_N=33, a
    converged solution of simple_FEM_1D_transient.m (2049 steps of dt = 1/2048, trapezoid rule, tau = 0.5302,
    L = 3.72374).  The similarity x -> c x, tau -> c tau, eta -> eta*2049/2048, phi -> phi*2049/2048 with
    c = sqrt(2048/2049) maps that march onto nsteps = 2049 of the engine (tests/test_oracle_golden.py)."""
    x, eta = fixtures["matlab33_x"], fixtures["matlab33_eta"]
    tau, L, c, s = 0.5302, 3.72374, np.sqrt(2048.0 / 2049.0), 2049.0 / 2048.0
    eng = sb.Engine(33, nsteps=2049, scheme=0, quadrature=sb.QUAD_TRAPEZOID, tau=tau * c, L=L * c)
    eng.residual(eta[1:-1] * s)
    phi = eng.phi()
    f0 = oracle.f0_given(x, tau)
    ref = oracle.residual(eta * s, f0, scheme=0, nsteps=2049, L=L * c, quadrature=1)
    assert np.abs(phi - ref["phi"]).max() < 1e-10 * np.abs(ref["phi"]).max()
    assert np.abs(eng.f0_given() - f0).max() < 1e-12
    assert np.abs(f0[1:-1] - phi[1:-1] * s).max() < 3e-7      # converged to the MATLAB run's 1e-7 (2.33e-7)
    eng.close()
