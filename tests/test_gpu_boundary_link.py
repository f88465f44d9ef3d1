"""Link proof of the drop-in boundary: the reference's OWN compiled solvers (oracle/_ref, built in place from
/root/reference by oracle/Makefile) are handed the RAW addresses of the library's residual callbacks — exactly
what a maintainer does when replacing SCFT_wrapper in `broydn(x, n, &check, SCFT_wrapper)` (drivescft.cc:301) or
`adm_chen(&SCFT_wrapper, ...)` (drivescft.cc:294-298).  No Python trampoline and no C adapter sits between the
reference code and the callback: ctypes passes the function address itself.

    broydn.c:44-46  + fdjac.c / lnsrch.c / qr*.c   <-  scftb_callback_nr1  (in[1..n], out[1..n])
    ADM_chen_C.c:18                                <-  scftb_callback_c0   (0-based)
"""
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


def addr(fn):
    return C.cast(fn, C.c_void_p).value


def test_nr1_callback_is_the_c0_callback_shifted_by_one(sb, oracle, fixtures):
    """scftb_callback_nr1 (NR 1-based arrays, the #define BROYDN default of drivescft.cc:218-243) against the
    0-based callback and the oracle."""
    N, n = 33, 2048
    L = sb.lib()
    em = fixtures["res32_eta"][1:-1]
    for scheme in (sb.IE_ROWSCALE, sb.IE_CONSISTENT, sb.IRK4_CONSISTENT):
        eng = sb.Engine(N, nsteps=n, scheme=scheme)
        eng.bind_global()
        xin = np.concatenate([[777.0], em])          # in[0] is never read by an NR callee
        out1 = np.full(N - 1, -555.0)
        L.scftb_callback_nr1(N - 2, xin.ctypes.data_as(_dp), out1.ctypes.data_as(_dp))
        out0 = np.zeros(N - 2)
        x0 = em.copy()
        L.scftb_callback_c0(N - 2, x0.ctypes.data_as(_dp), out0.ctypes.data_as(_dp))
        assert C.c_int.in_dll(L, "scftb_funcerr").value == 0
        assert out1[0] == -555.0 and np.array_equal(out1[1:], out0)
        x = oracle.mesh_uniform(N)
        ref = oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x), scheme=scheme, nsteps=n)
        assert np.abs(out0 - ref["out"]).max() < 1e-10
        eng.close()


def test_reference_broydn_drives_raw_nr1_callback(sb, oracle, fixtures):
    """the reference's broydn.c (sequential fdjac.c, its own qrdcmp/qrupdt/lnsrch) on scftb_callback_nr1 must land
    where scftb_broydn (same method, Jacobian as one device batch) lands on scftb_callback_c0"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    N, n, scheme = 33, 128, sb.IRK4_CONSISTENT
    L = sb.lib()
    x0 = fixtures["n33_eta"][1:-1] * 1.01
    eng = sb.Engine(N, nsteps=n, scheme=scheme, max_batch=N - 2)
    eng.bind_global()
    chk_r, x_r, err_r, jc_r = oracle.ref_broydn_raw(addr(L.scftb_callback_nr1), x0, 1e-9)
    assert C.c_int.in_dll(L, "scftb_funcerr").value == 0
    x = x0.copy()
    chk, err, jc = C.c_int(1), C.c_double(1e-9), C.c_int(0)
    rc = L.scftb_broydn(L.scftb_callback_c0, x.ctypes.data_as(_dp), N - 2, C.byref(chk), C.byref(err), C.byref(jc))
    assert rc == 0 and chk.value == 0 == chk_r
    assert jc.value == jc_r
    assert err.value < 1e-9 and err_r < 1e-9
    # same method, same residual: the iterates differ only through the summation order of the host algebra
    assert np.abs(x - x_r).max() < 1e-9 * np.abs(x_r).max()
    # broydn.c:215-228 hands back the iterate BEFORE the last accepted trial point together with the trial point's
    # residual norm (see SCFTB_BROYDN_KEEP_TRIAL in the header), so the returned field itself sits one step earlier
    out = eng.residual(x_r)
    assert np.abs(out).max() < 1e-7
    # and the device-resident variant agrees with the reference too
    rc, chk_d, x_d, err_d, jc_d = eng.broydn_device(x0, 1e-9)
    assert rc == 0 and chk_d == 0
    assert np.abs(x_d - x_r).max() < 1e-8 * np.abs(x_r).max()
    eng.close()


def test_reference_adm_chen_drives_raw_c0_callback(sb, oracle, fixtures, capfd):
    """the reference's adm_chen (ADM_chen_C.c) on the raw scftb_callback_c0: bit-identical to scftb_adm_chen and to the
    device-resident mixer, which see the same residual"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    N, n = 33, 128
    L = sb.lib()
    x0 = fixtures["res32_eta"][1:-1]
    cases = [(1e-30, 25, 0.99, 2), (1e-30, 40, 0.9, 15), (1e-3, 300, 0.9, 3)]
    # the reference exit(1)s on a NaN residual (ADM_chen_C.c:61-66): only schedules known not to diverge are run
    for scheme, sched in ((sb.IE_ROWSCALE, cases), (sb.IE_CONSISTENT, cases), (sb.IRK4_CONSISTENT, [(1e-30, 6, 0.9, 3)])):
        eng = sb.Engine(N, nsteps=n, scheme=scheme)
        eng.bind_global()
        for (tol, mi, lmd, nn) in sched:
            rc_r, x_r = oracle.ref_adm_chen_raw(addr(L.scftb_callback_c0), x0, tol, mi, lmd, nn)
            x_h = x0.copy()
            rc_h = L.scftb_adm_chen(L.scftb_callback_c0, x_h.ctypes.data_as(_dp), tol, mi, N - 2, lmd, nn, 0)
            rc_d, x_d, iters, err = eng.adm_chen_batch(x0, tol, mi, lmd, nn)
            assert rc_r == rc_h and (rc_h == 0) == (rc_d == 0)
            assert np.array_equal(x_r, x_h), (scheme, nn, np.abs(x_r - x_h).max())
            assert np.array_equal(x_r, x_d), (scheme, nn, np.abs(x_r - x_d).max())
        eng.close()
    capfd.readouterr()   # adm_chen prints its progress


def test_broydn_jacobian_uses_problem_0_on_a_sweep_engine(sb, fixtures):
    """The batched finite-difference Jacobian runs its n columns through the engine in one launch; on a sweep engine
    (every slot its own tau, L) all columns must still be evaluated with the parameters of problem 0 — the problem the
    callback solves — and the other slots' results must stay untouched."""
    N, n, scheme = 33, 128, sb.IE_CONSISTENT
    L = sb.lib()
    x0 = fixtures["n33_eta"][1:-1] * 1.01
    B = N - 2
    sweep_eng = sb.Engine(N, nsteps=n, scheme=scheme, max_batch=B)
    taus, Ls = np.linspace(0.40, 0.66, B), np.linspace(3.2, 4.2, B)
    taus[0], Ls[0] = 5.30252230020752e-01, 3.72374357332160      # problem 0: the reference's own (drivescft.cc:269)
    for p in range(B):
        sweep_eng.set_problem(p, taus[p], Ls[p])
    sweep_eng.residual(np.tile(x0, (B, 1)))
    phi_before = [sweep_eng.phi(p) for p in range(B)]
    plain = sb.Engine(N, nsteps=n, scheme=scheme, max_batch=B, tau=taus[0], L=Ls[0])
    res = []
    for eng in (sweep_eng, plain):
        eng.bind_global()
        x = x0.copy()
        chk, err, jc = C.c_int(1), C.c_double(1e-9), C.c_int(0)
        rc = L.scftb_broydn(L.scftb_callback_c0, x.ctypes.data_as(_dp), N - 2, C.byref(chk), C.byref(err), C.byref(jc))
        assert rc == 0 and chk.value == 0
        rc, chk_d, x_d, err_d, _ = eng.broydn_device(x0, 1e-9)
        assert rc == 0 and chk_d == 0
        res.append((x, x_d))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    for p in range(1, B):
        assert np.array_equal(sweep_eng.phi(p), phi_before[p])
    sweep_eng.close()
    plain.close()


def test_documented_drivescft_configuration_reproduces_the_reference_fixture(sb, oracle, fixtures):
    """INTEGRATION.md section 1 (drivescft.cc): SCFTB_IRK4_CONSISTENT, n = 2048, Romberg, sign +1, the reference's own
    broydn on scftb_callback_nr1.  Started from the reference's converged inputFiles/N=33_for_read.txt it must accept
    the field at once (header ERROR= 1.42e-9) and report the file's free energy 0.001945037280931."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    sec1 = doc.split("## 1.")[1].split("## 2.")[0]
    assert "scftb_config cfg = { SCFTB_IRK4_CONSISTENT" in sec1
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    N = 33
    L = sb.lib()
    eng = sb.Engine(N, nsteps=2048, scheme=sb.IRK4_CONSISTENT, quadrature=sb.QUAD_ROMBERG, sign=+1.0, max_batch=N - 2)
    eng.bind_global()
    x0 = fixtures["n33_eta"][1:-1]
    chk, x, err, jc = oracle.ref_broydn_raw(addr(L.scftb_callback_nr1), x0, 1e-8)
    assert chk == 0 and err < 2e-9
    assert np.array_equal(x, x0)
    eng.residual(x)
    assert eng.free_energy() == pytest.approx(float(fixtures["n33_F"]), abs=1e-15)
    # the implicit-Euler schemes are a different discretisation: the same field is NOT their fixed point
    e2 = sb.Engine(N, nsteps=2048, scheme=sb.IE_CONSISTENT)
    assert np.abs(e2.residual(x0)).max() > 1e-5
    e2.close()
    eng.close()
