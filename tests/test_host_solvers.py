"""Host-flow field-update solvers of the product library (scftb_adm_chen / scftb_adm / scftb_broydn)
driven through the C ABI with CPU toy callbacks, against (i) the outputs of the unmodified
reference C recorded in tests/golden/ref_outputs.npz and (ii) the oracle restatement.
No GPU work: the callbacks are plain Python functions."""
import ctypes as C

import numpy as np
import pytest

from toys import toyF, toy3, toy2, fp4

_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def L():
    import scft_b200
    return scft_b200.lib()


def wrap(pyfunc):
    from scft_b200.engine import FUNC

    def cb(n, pin, pout):
        x = np.ctypeslib.as_array(pin, shape=(n,)).copy()
        np.ctypeslib.as_array(pout, shape=(n,))[:] = pyfunc(x)
    return FUNC(cb)


def run_adm_chen(L, f, x0, tol, mi, lmd, nn, final=0):
    x = np.array(x0, dtype=np.float64)
    cb = wrap(f)
    rc = L.scftb_adm_chen(cb, x.ctypes.data_as(_dp), tol, mi, len(x), lmd, nn, final)
    return rc, x


def run_adm(L, f, x0, maxits=100000):
    x = np.array(x0, dtype=np.float64)
    chk = C.c_int(1)
    cb = wrap(f)
    rc = L.scftb_adm(cb, x.ctypes.data_as(_dp), len(x), C.byref(chk), maxits)
    return rc, chk.value, x


def run_broydn(L, f, x0, tolf, jc=0):
    x = np.array(x0, dtype=np.float64)
    chk, err, jcv = C.c_int(1), C.c_double(tolf), C.c_int(jc)
    cb = wrap(f)
    rc = L.scftb_broydn(cb, x.ctypes.data_as(_dp), len(x), C.byref(chk), C.byref(err), C.byref(jcv))
    return rc, chk.value, x, err.value, jcv.value


def test_library_exports_every_declared_symbol(L):
    import re, os
    import scft_b200
    hdr = open(os.path.join(os.path.dirname(scft_b200.__file__), "..", "include", "scft_b200.h")).read()
    declared = set(re.findall(r"\b(scftb(?:2d)?_[A-Za-z0-9_]+)\s*\(", hdr)) | {"scftb_funcerr"}
    declared -= {"scftb_func", "scftb2d_engine", "scftb2d_config"}
    assert declared == set(scft_b200.engine.EXPORTS)
    for s in sorted(declared):
        assert hasattr(L, s), s


def test_adm_chen_bitwise_vs_reference(L, refout):
    rc, x = run_adm_chen(L, toyF, [1., 2., 3.], 1e-13, 500, 0.9, 3)
    assert rc == 0 and np.array_equal(x, refout["admchen_toyF_x"])
    rc, x = run_adm_chen(L, toy3, [1., 2., 3.], 1e-12, 2000, 0.99, 30)
    assert rc == 0 and np.array_equal(x, refout["admchen_toy3_x"])


def test_adm_chen_iteration_limit_and_nan(L, oracle):
    rc, x = run_adm_chen(L, toy3, [1., 2., 3.], 1e-12, 5, 0.99, 30)
    rc2, x2, _, _ = oracle.adm_chen(toy3, [1., 2., 3.], 1e-12, 5, 0.99, 30)
    assert rc == 1 and rc2 == 1 and np.array_equal(x, x2)   # same field after the same iteration count
    rc, _ = run_adm_chen(L, lambda v: v * np.nan, [1., 2.], 1e-12, 5, 0.9, 3)
    assert rc == 3  # SCFTB_ERR_NAN instead of the reference's exit(1)


def test_adm_bitwise_vs_reference(L, refout):
    rc, chk, x = run_adm(L, toy2, [1., 2.])
    assert rc == 0 and chk == 0 and np.array_equal(x, refout["adm_toy2_x"])
    rc, chk, x = run_adm(L, fp4, [1., 2., 3., 4.])
    assert rc == 0 and chk == 0 and np.array_equal(x, refout["adm_fp4_x"])


def test_broydn_vs_reference(L, refout):
    rc, chk, x, err, jc = run_broydn(L, toyF, [1., 2., 3.], 1e-6)
    assert rc == 0 and chk == int(refout["broydn_toyF_check"]) and jc == int(refout["broydn_toyF_jc"])
    assert np.array_equal(x, refout["broydn_toyF_x"]) and err == float(refout["broydn_toyF_err"])
    rc, chk, x, err, jc = run_broydn(L, toy3, [1., 2., 3.], 1e-8)
    assert rc == 0 and chk == int(refout["broydn_toy3_check"])
    assert np.allclose(x, refout["broydn_toy3_x"], rtol=1e-12, atol=0)
    assert err == pytest.approx(float(refout["broydn_toy3_err"]), rel=1e-6)


def test_broydn_jacobian_reuse(L):
    rc, chk, x, err, jc = run_broydn(L, toyF, [1., 2., 3.], 1e-6)
    assert jc == 1
    # second call from a nearby point re-uses the stored QR (jc=1): no fdjac evaluations
    calls = []

    def counted(v):
        calls.append(1)
        return toyF(v)
    rc, chk, x2, err, jc = run_broydn(L, counted, x + 1e-3, 1e-6, jc=1)
    assert rc == 0 and chk == 0 and len(calls) < 3 + 6
    assert np.abs(toyF(x2)).max() < 1e-6
