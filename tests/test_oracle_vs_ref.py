"""Live comparison of the oracle with the UNMODIFIED reference C compiled into oracle/_ref
(romint, polint, spline_chen+gaussj, gaussj, adm_chen, adm, broydn).  Skipped when the prebuilt
libraries are absent (they are git-ignored; build() makes them when /root/reference exists)."""
import numpy as np
import pytest

from oracle import oracle as O
from toys import toyF, toy3, fp4

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def test_romint_live():
    rng = np.random.default_rng(7)
    for m in (16, 256, 2048, 65536):
        f = rng.standard_normal(m + 1)
        assert O.romint(f, 0.37 / m) == O.ref_romint(f, 0.37 / m)


def test_spline_live_dense_gaussj_vs_banded():
    rng = np.random.default_rng(8)
    for N in (9, 33, 65, 257):
        x = O.mesh_uniform(N)
        y = rng.standard_normal(N - 2) * 10
        a, b = O.spline(x[1:-1], y, x), O.ref_spline(x[1:-1], y, x)
        assert np.abs(a - b).max() < 1e-12 * np.abs(y).max()


def test_gaussj_live():
    rng = np.random.default_rng(9)
    for n in (1, 2, 5, 30, 50):
        A, B = rng.standard_normal((n, n)), rng.standard_normal((n, 1))
        r1, r2 = O.gaussj(A, B), O.ref_gaussj(A, B)
        assert r1[0] == r2[0] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])


def test_adm_chen_live(capfd):
    for (f, x0, tol, mi, lmd, nn) in [(toyF, [1., 2., 3.], 1e-13, 500, 0.9, 3), (toy3, [1., 2., 3.], 1e-12, 2000, 0.99, 30),
                                      (toyF, [0.3, -0.2, 0.9], 1e-10, 40, 0.5, 2)]:
        rc, x, _, _ = O.adm_chen(f, x0, tol, mi, lmd, nn)
        rc2, x2 = O.ref_adm_chen(f, x0, tol, mi, lmd, nn)
        assert rc == rc2 and np.array_equal(x, x2)


def test_adm_live(capfd):
    rc, x, _, _ = O.adm(fp4, [1., 2., 3., 4.])
    rc2, x2 = O.ref_adm(fp4, [1., 2., 3., 4.])
    assert rc == rc2 == 0 and np.array_equal(x, x2)


def test_reference_broydn_converges_on_toys(capfd):
    chk, x, err, jc = O.ref_broydn(toyF, [1., 2., 3.], 1e-6)
    assert chk == 0 and jc == 1 and np.abs(toyF(x)).max() < 1e-6
