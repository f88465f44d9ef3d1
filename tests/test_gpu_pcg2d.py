"""2-D path on the GPU (CSR assembly + Jacobi-PCG per contour step) against the numpy/scipy oracle
(oracle/oracle2d.py: same Q1 matrices, sparse LU per field update) and against the 1-D engine on
y-invariant fields."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
REL = 1e-10


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def field(nx, ny, L, rng, kind):
    x = L * np.arange(nx + 1) / nx
    y = np.arange(ny + 1) / max(ny, 1)
    base = 3.0 * np.cos(2 * np.pi * x / L)[:, None] * np.ones((1, ny + 1))
    if kind == "yinv":
        return base.ravel()
    if kind == "ymod":
        return (base * (1 + 0.1 * np.cos(2 * np.pi * y)[None, :])).ravel()
    return (base + rng.standard_normal((nx + 1, ny + 1))).ravel()


@pytest.mark.parametrize("nx,ny,nsteps", [(16, 3, 16), (32, 8, 32), (40, 17, 16), (64, 64, 16)])
def test_csr_matrices_match_oracle(sb, nx, ny, nsteps):
    from oracle import oracle as O, oracle2d as O2
    rng = np.random.default_rng(nx)
    L, Ly = O.L_REF, O.L_REF * ny / nx * 1.3
    eta = field(nx, ny, L, rng, "rand")
    eng = sb.Engine2D(nx, ny, L=L, Ly=Ly, nsteps=nsteps, maxit=1)
    eng.residual(eta)
    rowptr, col, vt, va = eng.csr()
    nd = (nx + 1) * (ny + 1)
    T = sp.csr_matrix((vt, col, rowptr), shape=(nd, nd))
    A = sp.csr_matrix((va, col, rowptr), shape=(nd, nd))
    Tr, Ar = O2.system_matrices(nx, ny, L, Ly, eta, nsteps)
    assert abs(T - Tr).max() < 1e-14 * abs(Tr).max()
    assert abs(A - Ar).max() < 1e-14 * abs(Ar).max()
    assert abs(T - T.T).max() < 1e-15          # SPD form: CG applies
    assert rowptr[-1] <= 9 * nd
    eng.close()


@pytest.mark.parametrize("nx,ny,nsteps,kind", [(32, 4, 64, "yinv"), (32, 8, 64, "ymod"), (48, 12, 32, "rand"), (128, 16, 16, "rand"),
                                               (128, 128, 16, "ymod")])   # y-modulated field on a 129^2 mesh
def test_march_matches_oracle(sb, nx, ny, nsteps, kind):
    from oracle import oracle as O, oracle2d as O2
    rng = np.random.default_rng(ny)
    L, Ly = O.L_REF, O.L_REF * ny / nx
    eta = field(nx, ny, L, rng, kind)
    eng = sb.Engine2D(nx, ny, L=L, Ly=Ly, nsteps=nsteps, rtol=1e-13)
    out = eng.residual(eta)
    ref = O2.residual(nx, ny, L, Ly, eta, nsteps=nsteps)
    scale = np.abs(ref["phi"]).max()
    assert np.abs(eng.phi() - ref["phi"]).max() < REL * scale
    assert np.abs(out - ref["out"]).max() < REL * scale
    it, ms = eng.stats()
    assert it > nsteps
    eng.close()


def test_y_invariant_field_equals_1d_engine(sb, fixtures):
    """with eta(x,y) = eta(x) the 2-D solution must equal the 1-D one"""
    from oracle import oracle as O
    nx, ny, n = 32, 6, 128
    x = O.mesh_uniform(nx + 1)
    ef = O.eta_full(x, fixtures["n33_eta"][1:-1])
    e1 = sb.Engine(nx + 1, nsteps=n, scheme=sb.IE_CONSISTENT)
    e1.residual(fixtures["n33_eta"][1:-1])
    phi1 = e1.phi()
    e2 = sb.Engine2D(nx, ny, nsteps=n, rtol=1e-13)
    e2.residual(np.repeat(ef, ny + 1))
    phi2 = e2.phi().reshape(nx + 1, ny + 1)
    assert np.abs(phi2 - phi2[:, [0]]).max() < 1e-11
    assert np.abs(phi2[:, 0] - phi1).max() < REL
    e1.close(); e2.close()


def test_y_invariant_field_equals_1d_engine_at_config_scale(sb, fixtures):
    """BASELINE.json configs[3] scale: 1024 x 1023 cells (1.05M DOFs), n = 2048 implicit-Euler steps, y-invariant field from the
    m=1024 spectral result: the 2-D density must equal the 1-D engine's (IE on the consistent matrices) to 1e-10 on every node"""
    nx, ny, n = 1024, 1023, 2048
    e1 = sb.Engine(nx + 1, nsteps=n, scheme=sb.IE_CONSISTENT)
    e1.residual(fixtures["res1024_eta"][1:-1])
    phi1, ef = e1.phi(), e1.eta_full()
    e1.close()
    e2 = sb.Engine2D(nx, ny, nsteps=n, rtol=1e-12)
    out = e2.residual(np.repeat(ef, ny + 1))
    phi2 = e2.phi().reshape(nx + 1, ny + 1)
    it, ms = e2.stats()
    e2.close()
    assert it >= n
    assert np.abs(phi2 - phi2[:, [0]]).max() < 1e-11
    assert np.abs(phi2[:, 0] - phi1).max() < REL
    assert np.isfinite(out).all()
