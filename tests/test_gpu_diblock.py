"""Two-species (AB diblock) march on the GPU — q and q+ as separate sweeps, every q slice kept in HBM — against
oracle/scft_oracle.c::orc_residual_ab (not in the reference: parity pinned by oracle only).
Tolerances: phi_A, phi_B, Q relative 1e-10; residual absolute 1e-10 * max(1, |eta|, chiN)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-10
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def _fields(rng, N, scale=2.0):
    return rng.standard_normal(N - 2) * scale, rng.standard_normal(N - 2) * scale


def _check(eng, p, w, ref, chi):
    ni = eng.ni
    pa, pb = eng.phi_ab(p)
    assert np.abs(pa - ref["phiA"]).max() <= REL * np.abs(ref["phiA"]).max()
    assert np.abs(pb - ref["phiB"]).max() <= REL * np.abs(ref["phiB"]).max()
    assert abs(eng.Q(p) - ref["Q"]) <= REL * ref["Q"]
    return REL * max(1.0, np.abs(w).max(), abs(chi))


@pytest.mark.parametrize("N", [33, 65, 129, 257, 513, 1025, 2049])
@pytest.mark.parametrize("scheme", [0, 1])
def test_parity_every_kernel_shape(sb, N, scheme):
    n, jf, chi = 128, 48, 9.0
    rng = np.random.default_rng(N + scheme)
    a, b = _fields(rng, N)
    x = O.mesh_uniform(N)
    eng = sb.Engine(N, nsteps=n, scheme=scheme)
    eng.set_diblock(jf / n, chi)
    w = np.concatenate([a, b])
    out = eng.residual_ab(w)
    ref = O.residual_ab(O.eta_full(x, a), O.eta_full(x, b), jf, chi, O.f0_given(x), scheme=scheme, nsteps=n)
    tol = _check(eng, 0, w, ref, chi)
    assert np.abs(out - ref["out"]).max() <= tol
    eng.close()


@pytest.mark.parametrize("quad,n,jf", [(0, 2048, 1024), (0, 2048, 512), (1, 2048, 700), (0, 256, 64), (1, 33, 16), (0, 64, 1),
                                       (0, 64, 63)])
def test_parity_block_quadratures(sb, fixtures, quad, n, jf):
    """Romberg on blocks of 2^k >= 16 steps, trapezoid otherwise; odd step counts; one-step blocks"""
    N, chi = 33, 4.0
    em = fixtures["n33_eta"][1:-1]
    a, b = em * 1.05, em * 0.9 + 0.2
    x = O.mesh_uniform(N)
    eng = sb.Engine(N, nsteps=n, scheme=0, quadrature=quad)
    eng.set_diblock(jf / n, chi)
    w = np.concatenate([a, b])
    out = eng.residual_ab(w)
    ref = O.residual_ab(O.eta_full(x, a), O.eta_full(x, b), jf, chi, O.f0_given(x), scheme=0, nsteps=n, quadrature=quad)
    tol = _check(eng, 0, w, ref, chi)
    assert np.abs(out - ref["out"]).max() <= tol
    eng.close()


def test_equal_fields_give_the_one_sweep_residual(sb, fixtures):
    """with q+ marched separately the density must match the one-sweep form"""
    N, n = 129, 2048
    rng = np.random.default_rng(3)
    em = rng.standard_normal(N - 2)
    eng = sb.Engine(N, nsteps=n, scheme=1, quadrature=1)
    one = eng.residual(em)
    phi = eng.phi()
    eng.set_diblock(700 / n, 6.0)
    out = eng.residual_ab(np.concatenate([em, em]))
    pa, pb = eng.phi_ab()
    assert np.abs(pa + pb - phi).max() < 1e-12
    assert np.abs(out[: N - 2] - one).max() < 1e-12
    eng.close()


def test_batch_with_distinct_parameters(sb):
    """a chi N x (tau, L) sweep: every problem equals its own oracle evaluation"""
    N, n, B, jf = 129, 64, 21, 24
    rng = np.random.default_rng(17)
    eng = sb.Engine(N, nsteps=n, scheme=0, max_batch=B)
    taus, Ls, chis = np.linspace(0.40, 0.66, B), np.linspace(3.2, 4.2, B), np.linspace(0.0, 20.0, B)
    w = rng.standard_normal((B, 2 * (N - 2))) * 2
    for p in range(B):
        eng.set_problem(p, taus[p], Ls[p])
        eng.set_diblock(jf / n, chis[p], p=p)
    out = eng.residual_ab(w)
    for p in range(B):
        x = O.mesh_uniform(N, Ls[p])
        ref = O.residual_ab(O.eta_full(x, w[p, : N - 2]), O.eta_full(x, w[p, N - 2:]), jf, chis[p], O.f0_given(x, taus[p]),
                            scheme=0, nsteps=n, L=Ls[p])
        tol = _check(eng, p, w[p], ref, chis[p])
        assert np.abs(out[p] - ref["out"]).max() <= tol
    eng.close()


def test_more_problems_than_resident_slots(sb):
    """problems are strided over the resident CTAs; each slot's history is reused"""
    N, n, B, jf = 33, 32, 2500, 8
    rng = np.random.default_rng(23)
    eng = sb.Engine(N, nsteps=n, scheme=0, max_batch=B)
    eng.set_diblock(jf / n, 3.0)
    w = rng.standard_normal((B, 2 * (N - 2)))
    out = eng.residual_ab(w)
    x = O.mesh_uniform(N)
    f0 = O.f0_given(x)
    for p in (0, 1, 1183, 1184, 2499):
        ref = O.residual_ab(O.eta_full(x, w[p, : N - 2]), O.eta_full(x, w[p, N - 2:]), jf, 3.0, f0, scheme=0, nsteps=n)
        assert np.abs(out[p] - ref["out"]).max() <= 1e-10 * max(1.0, np.abs(w[p]).max())
    eng.close()


def test_nonuniform_mesh(sb, fixtures):
    x = fixtures["matlab43_x"]
    N, n, jf, chi = len(x), 64, 40, 5.0
    L = x[-1] - x[0]
    rng = np.random.default_rng(31)
    a, b = _fields(rng, N)
    for scheme in (0, 1):
        eng = sb.Engine(N, nsteps=n, scheme=scheme, tau=0.5302, L=L, x=x)
        eng.set_diblock(jf / n, chi)
        w = np.concatenate([a, b])
        out = eng.residual_ab(w)
        ref = O.residual_ab(O.eta_full(x, a), O.eta_full(x, b), jf, chi, O.f0_given(x, 0.5302), scheme=scheme, nsteps=n,
                            L=L, x=x)
        tol = _check(eng, 0, w, ref, chi)
        assert np.abs(out - ref["out"]).max() <= tol
        eng.close()


def test_broyden_converges_a_diblock_film(sb, fixtures):
    """chi N continuation 0 -> 5 -> 12 at fA = 1/4 with the host-flow broydn on the two-species callback (Jacobian columns
    as device batches); the oracle confirms the converged fields"""
    N, n, jf = 33, 256, 64
    ni = N - 2
    em = fixtures["n33_eta"][1:-1]
    eng = sb.Engine(N, nsteps=n, scheme=0, max_batch=2 * ni)
    eng.bind_global()
    L = sb.lib()
    x = np.concatenate([em, em])
    xm = O.mesh_uniform(N)
    f0 = O.f0_given(xm)
    for chi in (0.0, 5.0, 12.0):
        eng.set_diblock(jf / n, chi)
        chk, err, jc = C.c_int(1), C.c_double(1e-9), C.c_int(0)
        rc = L.scftb_broydn(L.scftb_callback_ab_c0, x.ctypes.data_as(_dp), 2 * ni, C.byref(chk), C.byref(err), C.byref(jc))
        assert rc == 0 and chk.value == 0 and err.value < 1e-9
        ref = O.residual_ab(O.eta_full(xm, x[:ni]), O.eta_full(xm, x[ni:]), jf, chi, f0, scheme=0, nsteps=n)
        assert np.abs(ref["out"]).max() < 1e-7
    assert np.abs(x[:ni] - x[ni:]).max() > 1.0          # a genuinely two-species solution
    eng.close()


def test_argument_errors(sb):
    eng = sb.Engine(33, nsteps=64, scheme=0)
    with pytest.raises(sb.ScftError):
        eng.residual_ab(np.zeros(62))                   # set_diblock not called
    with pytest.raises(sb.ScftError):
        eng.set_diblock(0.3, 1.0)                       # 0.3 * 64 is not an integer
    with pytest.raises(sb.ScftError):
        eng.set_diblock(1.0, 1.0)
    eng.close()
    eng = sb.Engine(33, nsteps=64, scheme=2)
    with pytest.raises(sb.ScftError):
        eng.set_diblock(0.5, 1.0)                       # IRK4: implicit-Euler schemes only
    eng.close()


def test_device_mixers_on_two_species_engine_equal_host_flow_bitwise(sb, fixtures):
    """adm_chen and adm, device-resident, on the (eta_A, eta_B) vector of a diblock engine: the same iterates, bit for bit,
    as the reference-shaped host flows driven by scftb_callback_ab_c0 / its fixed-point image"""
    N, n, jf = 33, 128, 32
    ni = N - 2
    em = fixtures["n33_eta"][1:-1]
    x0 = np.concatenate([em, 0.8 * em])
    L = sb.lib()
    eng = sb.Engine(N, nsteps=n, scheme=0, max_batch=2)
    eng.set_diblock(jf / n, 3.0)
    eng.bind_global()
    for (tol, mi, lmd, nn) in [(1e-30, 6, 0.9, 3), (1e-30, 40, 0.99, 2), (1e-30, 45, 0.9, 15)]:
        xh = x0.copy()
        rc_h = L.scftb_adm_chen(L.scftb_callback_ab_c0, xh.ctypes.data_as(_dp), tol, mi, 2 * ni, lmd, nn, 0)
        rc_d, xd, iters, err = eng.adm_chen_batch(np.stack([x0, 1.01 * x0]), tol, mi, lmd, nn)
        assert (rc_h == 0) == (rc_d == 0)
        assert xd.shape == (2, 2 * ni) and np.array_equal(xh, xd[0]), np.abs(xh - xd[0]).max()
    # adm: x -> x + F(x) on the two-species residual
    FUNC = C.CFUNCTYPE(None, C.c_int, _dp, _dp)

    @FUNC
    def fixed_point(nn_, pin, pout):
        L.scftb_callback_ab_c0(nn_, pin, pout)
        for i in range(nn_):
            pout[i] += pin[i]

    for maxits in (1, 3, 14):
        xh = x0.copy()
        chk = C.c_int(1)
        L.scftb_adm(fixed_point, xh.ctypes.data_as(_dp), 2 * ni, C.byref(chk), maxits)
        rc_d, xd, iters, err = eng.adm_batch(x0[None, :], maxits)
        assert np.array_equal(xh, xd[0]), (maxits, np.abs(xh - xd[0]).max())
    eng.close()
