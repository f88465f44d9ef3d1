"""Device-resident Broyden (scftb_broydn_device) against the host flow scftb_broydn, which reproduces
the reference's broydn.c bit for bit (tests/test_host_solvers.py), both on the GPU residual."""
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


def host_broydn(sb, eng, x0, tolf):
    L = sb.lib()
    eng.bind_global()
    x = np.array(x0, dtype=np.float64)
    chk, err, jc = C.c_int(1), C.c_double(tolf), C.c_int(0)
    rc = L.scftb_broydn(L.scftb_callback_c0, x.ctypes.data_as(_dp), len(x), C.byref(chk), C.byref(err), C.byref(jc))
    return rc, chk.value, x, err.value, jc.value


@pytest.mark.parametrize("N,scheme,nsteps,scale,tolf", [(33, 1, 256, 1.02, 1e-8), (33, 2, 256, 1.01, 1e-10), (65, 0, 128, 1.0, 1e-10),
                                                        (129, 1, 64, 1.0, 1e-10), (33, 1, 128, 1.01, 1e-10)])
def test_device_broyden_matches_host_flow(sb, oracle, fixtures, N, scheme, nsteps, scale, tolf):
    """same status (rc, check, jc), same achieved error and the same field as the host flow — including the
    cases where the line search gives up (check = 1), which the reference logic reports the same way"""
    x = oracle.mesh_uniform(33)
    em = fixtures["n33_eta"][1:-1] * scale
    Nc = 33
    while Nc < N:                                   # refine the N=33 field up to the requested mesh
        x, em = sb.refine_mesh(x, em)
        Nc = 2 * Nc - 1
    eng = sb.Engine(N, nsteps=nsteps, scheme=scheme, max_batch=N - 2)
    rc_h, chk_h, x_h, err_h, jc_h = host_broydn(sb, eng, em, tolf)
    rc_d, chk_d, x_d, err_d, jc_d = eng.broydn_device(em, tolf)
    assert (rc_h, chk_h, jc_h) == (rc_d, chk_d, jc_d)
    assert abs(err_d - err_h) < 1e-3 * err_h + 1e-13
    assert np.abs(x_d - x_h).max() < 1e-9 * np.abs(x_h).max()
    r_h, r_d = np.abs(eng.residual(x_h)).max(), np.abs(eng.residual(x_d)).max()
    assert abs(r_h - r_d) < 1e-2 * r_h + 1e-12 and r_d < 1e-7
    eng.close()


def test_device_broyden_converges_and_batches_the_jacobian(sb, oracle, fixtures):
    N = 33
    eng = sb.Engine(N, nsteps=128, scheme=1, max_batch=N - 2)
    em = fixtures["n33_eta"][1:-1] * 1.01
    before = sb.launch_count()
    rc, chk, x, err, jc = eng.broydn_device(em, 1e-8)
    used = sb.launch_count() - before
    assert rc == 0 and chk == 0 and jc == 1 and err < 1e-8
    assert used < 400      # one batched Jacobian launch, not n
    assert np.abs(eng.residual(x)).max() < 1e-7
    # second solve from a nearby point re-uses the stored device QR (jc = 1)
    rc, chk, x2, err2, jc2 = eng.broydn_device(x + 1e-4, 1e-8, jc=1)
    assert rc == 0 and np.abs(x2 - x).max() < 1e-5
    eng.close()


@pytest.mark.parametrize("N,scale", [(33, 1.02), (65, 1.0), (129, 1.0)])
def test_keep_trial_returns_the_field_whose_residual_is_reported(sb, oracle, fixtures, N, scale):
    """SCFTB_BROYDN_KEEP_TRIAL: *err is the residual norm of the returned x (the reference flow may hand back the
    previous iterate together with the trial's norm, broydn.c:215-228 / lnsrch.c)"""
    x = oracle.mesh_uniform(33)
    em = fixtures["n33_eta"][1:-1] * scale
    Nc = 33
    while Nc < N:
        x, em = sb.refine_mesh(x, em)
        Nc = 2 * Nc - 1
    eng = sb.Engine(N, nsteps=2048, scheme=0, max_batch=N - 2)
    rc, chk, xs, err, jc = eng.broydn_device(em, 1e-9, keep_trial=True)
    assert rc == 0 and chk == 0 and err < 1e-9
    assert abs(np.abs(eng.residual(xs)).max() - err) <= 1e-12
    eng.close()


def test_device_broyden_is_reproducible(sb, oracle, fixtures):
    """the cooperative QR must not depend on block timing: repeated solves give the same bits"""
    N = 257
    x = oracle.mesh_uniform(33)
    em = fixtures["n33_eta"][1:-1]
    Nc = 33
    while Nc < N:
        x, em = sb.refine_mesh(x, em)
        Nc = 2 * Nc - 1
    eng = sb.Engine(N, nsteps=256, scheme=0, max_batch=N - 2)
    runs = [eng.broydn_device(em, 1e-10) for _ in range(4)]
    for r in runs[1:]:
        assert r[1] == runs[0][1] and r[3] == runs[0][3] and np.array_equal(r[2], runs[0][2])
    eng.close()
