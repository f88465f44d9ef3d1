"""Error behaviour and edge cases of the C ABI on the GPU box: argument validation, size limits,
NaN propagation (the reference exit(1)s, ADM_chen_C.c:61-66; the library returns a status)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def test_argument_validation(sb):
    with pytest.raises(sb.ScftError, match="Romberg"):
        sb.Engine(33, nsteps=100)                      # romint needs n = 2^k >= 16 (romint.c:28-33)
    with pytest.raises(sb.ScftError, match="N too large"):
        sb.Engine(5000, nsteps=16)                     # register-resident march: N <= 4098
    with pytest.raises(sb.ScftError):
        sb.Engine(3, nsteps=16)
    with pytest.raises(sb.ScftError, match="unknown scheme"):
        sb.Engine(33, nsteps=16, scheme=7)
    eng = sb.Engine(33, nsteps=16, max_batch=2)
    with pytest.raises(sb.ScftError, match="nprob"):
        eng.residual(np.zeros((3, 31)))                # more problems than the engine holds
    with pytest.raises(sb.ScftError, match="store_history"):
        eng.q_history()
    eng.close()


def test_smallest_and_largest_meshes(sb, oracle):
    for N, n in [(4, 16), (5, 16), (34, 16), (4098, 16)]:
        rng = np.random.default_rng(N)
        x = oracle.mesh_uniform(N)
        em = rng.standard_normal(N - 2)
        eng = sb.Engine(N, nsteps=n, scheme=1)
        eng.residual(em)
        ref = oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x), scheme=1, nsteps=n)
        assert np.abs(eng.phi() - ref["phi"]).max() < 1e-10 * np.abs(ref["phi"]).max()
        eng.close()


def test_nan_field_is_reported_not_fatal(sb, fixtures):
    eng = sb.Engine(33, nsteps=16)
    em = fixtures["res32_eta"][1:-1].copy()
    em[7] = np.nan
    out = eng.residual(em)
    assert np.isnan(out).any()
    with pytest.raises(sb.ScftError, match="NaN"):      # SCFTB_ERR_NAN
        eng.adm_chen_batch(em, 1e-6, 10, 0.9, 3)
    eng.bind_global()
    L = sb.lib()
    x = em.copy()
    assert L.scftb_adm_chen(L.scftb_callback_c0, x.ctypes.data_as(_dp), 1e-6, 10, 31, 0.9, 3, 0) == 3
    eng.close()


def test_callback_without_bound_engine_sets_funcerr(sb):
    L = sb.lib()
    L.scftb_bind_global(None)
    x, y = np.zeros(31), np.zeros(31)
    L.scftb_callback_c0(31, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp))
    assert C.c_int.in_dll(L, "scftb_funcerr").value == 1
    eng = sb.Engine(33, nsteps=16)
    eng.bind_global()
    assert C.c_int.in_dll(L, "scftb_funcerr").value == 0
    eng.close()


def test_mixer_window_limits(sb):
    eng = sb.Engine(33, nsteps=16)
    with pytest.raises(sb.ScftError, match="window"):
        sb.AndersonBatch(eng, 1, nn=51)
    m = sb.AndersonBatch(eng, 1, nn=0)                 # nn = 0: pure relaxation X += (1-lk) Y
    m.reset(np.zeros(31))
    m.iterate_device(0)
    done, iters, err = m.status(0)
    assert err[0] > 0
    m.close(); eng.close()


def test_odd_number_of_steps_with_trapezoid(sb, oracle, fixtures):
    """no unpaired middle slice when n is odd: every slice j > n/2 pairs with n-j < n/2"""
    N = 33
    x = oracle.mesh_uniform(N)
    em = fixtures["res32_eta"][1:-1]
    for n in (33, 7, 3):
        for scheme in (0, 2):
            eng = sb.Engine(N, nsteps=n, scheme=scheme, quadrature=sb.QUAD_TRAPEZOID)
            eng.residual(em)
            ref = oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x), scheme=scheme, nsteps=n, quadrature=1)
            assert np.abs(eng.phi() - ref["phi"]).max() < 1e-10 * np.abs(ref["phi"]).max()
            eng.close()


def test_large_host_batch_goes_through_the_chunked_copy_pipeline(sb, oracle):
    """scftb_residual_batch splits batches of >= 4 waves of resident CTAs into chunks whose copies overlap the march:
    every chunk must use its own problems' parameters (tau, L), outputs and history"""
    N, n, B = 33, 32, 12000
    rng = np.random.default_rng(77)
    eng = sb.Engine(N, nsteps=n, scheme=0, max_batch=B, store_history=True)
    taus, Ls = rng.uniform(0.40, 0.66, B), rng.uniform(3.2, 4.2, B)
    for p in range(B):
        eng.set_problem(p, taus[p], Ls[p])
    eta = rng.standard_normal((B, N - 2))
    out = eng.residual(eta)
    for p in (0, 1, 2999, 3000, 5999, 6001, 9000, 11999):
        x = oracle.mesh_uniform(N, Ls[p])
        ref = oracle.residual(oracle.eta_full(x, eta[p]), oracle.f0_given(x, taus[p]), scheme=0, nsteps=n, L=Ls[p],
                              want_hist=True)
        assert np.abs(out[p] - ref["out"]).max() < 1e-12
        assert np.abs(eng.phi(p) - ref["phi"]).max() < 1e-12
        assert abs(eng.Q(p) - ref["Q"]) < 1e-12
        assert np.abs(eng.q_history(p) - ref["hist"]).max() < 1e-12
    eng.close()


def test_large_host_batch_on_a_nonuniform_mesh(sb, oracle, fixtures):
    x = fixtures["matlab43_x"]
    N, n, B = len(x), 16, 8000
    L = x[-1] - x[0]
    rng = np.random.default_rng(78)
    eng = sb.Engine(N, nsteps=n, scheme=1, max_batch=B, tau=0.5302, L=L, x=x)
    eta = rng.standard_normal((B, N - 2))
    out = eng.residual(eta)
    f0 = oracle.f0_given(x, 0.5302)
    for p in (0, 2500, 5000, 7999):
        ref = oracle.residual(oracle.eta_full(x, eta[p]), f0, scheme=1, nsteps=n, L=L, x=x)
        assert np.abs(out[p] - ref["out"]).max() < 1e-12
        assert np.abs(eng.eta_full(p) - oracle.eta_full(x, eta[p])).max() < 1e-12
    eng.close()
