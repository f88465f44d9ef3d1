"""Two-species (AB diblock) restatement in the oracle (orc_residual_ab; not in the reference):
pinned by its homopolymer limit against the one-sweep residual, by the A<->B mirror symmetry and by an
independent dense numpy march."""
import numpy as np
import pytest

from oracle import oracle as O


def _fields(fixtures, N=33, L=O.L_REF):
    em = fixtures["n33_eta"][1:-1]
    x = O.mesh_uniform(N, L)
    return x, O.f0_given(x), O.eta_full(x, em)


@pytest.mark.parametrize("scheme", [O.IE_ROWSCALE, O.IE_CONSISTENT])
@pytest.mark.parametrize("quad,jf", [(O.QUAD_TRAPEZOID, 700), (O.QUAD_TRAPEZOID, 512), (O.QUAD_ROMBERG, 1024)])
def test_equal_fields_reduce_to_the_one_sweep_residual(fixtures, scheme, quad, jf):
    x, f0, ef = _fields(fixtures)
    one = O.residual(ef, f0, scheme=scheme, nsteps=2048, quadrature=quad)
    ab = O.residual_ab(ef, ef, jf, 7.0, f0, scheme=scheme, nsteps=2048, quadrature=quad)
    ni = len(ef) - 2
    assert np.abs(ab["phiA"] + ab["phiB"] - one["phi"]).max() < 1e-13
    assert np.abs(ab["out"][:ni] - one["out"]).max() < 1e-13
    assert abs(ab["Q"] - one["Q"]) < 1e-14
    # exchange residual: eta_A - eta_B = 0, so only -chiN*(phiB - phiA) remains
    assert np.abs(ab["out"][ni:] + 7.0 * (ab["phiB"] - ab["phiA"])[1:-1]).max() < 1e-15


def test_block_mirror_symmetry(fixtures):
    """relabelling the blocks (A<->B, f -> 1-f) swaps the densities: the chain read from its other end"""
    x, f0, ef = _fields(fixtures)
    a, b = ef * 1.1, ef * 0.8 + 0.3
    r1 = O.residual_ab(a, b, 600, 5.0, f0, scheme=O.IE_ROWSCALE, nsteps=2048, quadrature=O.QUAD_TRAPEZOID)
    r2 = O.residual_ab(b, a, 2048 - 600, 5.0, f0, scheme=O.IE_ROWSCALE, nsteps=2048, quadrature=O.QUAD_TRAPEZOID)
    assert np.abs(r1["phiA"] - r2["phiB"]).max() < 1e-13 and np.abs(r1["phiB"] - r2["phiA"]).max() < 1e-13


def test_against_dense_numpy_march():
    """independent restatement: dense matrices, numpy.linalg.solve, explicit trapezoid sums"""
    N, n, jf, L, chi = 12, 40, 15, 2.0, 3.0
    rng = np.random.default_rng(5)
    x = O.mesh_uniform(N, L)
    ea, eb = rng.standard_normal(N) * 3, rng.standard_normal(N) * 3
    f0 = O.f0_given(x, 0.5)
    h, dt, ni = L / (N - 1), 1.0 / n, N - 2
    A = (np.diag(np.full(ni, 4 * h / 6)) + np.diag(np.full(ni - 1, h / 6), 1) + np.diag(np.full(ni - 1, h / 6), -1))
    B = (np.diag(np.full(ni, 2 / h)) + np.diag(np.full(ni - 1, -1 / h), 1) + np.diag(np.full(ni - 1, -1 / h), -1))
    TA = A + dt * (B + np.diag(ea[1:-1]) @ A)          # row-scaled C (1D_FEM.c:104-105)
    TB = A + dt * (B + np.diag(eb[1:-1]) @ A)
    q = np.ones(ni); hq = [q.copy()]
    for s in range(1, n + 1):
        q = np.linalg.solve(TA if s <= jf else TB, A @ q); hq.append(q.copy())
    d = np.ones(ni); hd = [d.copy()]
    for s in range(1, n + 1):
        d = np.linalg.solve(TB if s <= n - jf else TA, A @ d); hd.append(d.copy())
    v = np.array([hq[j] * hd[n - j] for j in range(n + 1)])
    trap = lambda y: dt * (y[0] / 2 + y[1:-1].sum(axis=0) + y[-1] / 2)
    pa, pb = trap(v[: jf + 1]), trap(v[jf:])
    r = O.residual_ab(ea, eb, jf, chi, f0, scheme=O.IE_ROWSCALE, nsteps=n, L=L, quadrature=O.QUAD_TRAPEZOID)
    assert np.abs(r["phiA"][1:-1] - pa).max() < 1e-13 and np.abs(r["phiB"][1:-1] - pb).max() < 1e-13
    assert np.abs(r["out"][:ni] - (f0[1:-1] - pa - pb)).max() < 1e-13
    assert np.abs(r["out"][ni:] - (ea[1:-1] - eb[1:-1] - chi * (pb - pa))).max() < 1e-12
    assert abs(r["Q"] - h * hq[-1].sum() / L) < 1e-14


def test_rejects_irk4_and_bad_block(fixtures):
    x, f0, ef = _fields(fixtures)
    with pytest.raises(ValueError):
        O.residual_ab(ef, ef, 100, 0.0, f0, scheme=O.IRK4_CONSISTENT, nsteps=2048)
    with pytest.raises(ValueError):
        O.residual_ab(ef, ef, 2048, 0.0, f0, scheme=O.IE_ROWSCALE, nsteps=2048)
