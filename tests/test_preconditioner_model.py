"""The design claim behind the preconditioned mixer (scft_b200/csrc/pmixer.cu), checked on the CPU oracle:
the inverse of the residual Jacobian dF/d eta, F = phi0 - phi, is local —  inv(J) = I + Psi^-1 (1/2)(-Lap_h) Psi^-1 + (small,
smooth), Psi = diag(sqrt(phi)) — at the reference's own converged N = 33 field (IRK4, DEALII_SCFT/inputFiles/N=33_for_read.txt)."""
import numpy as np


def _jacobian(oracle, x, f0, em, scheme, nsteps):
    F0 = oracle.residual(oracle.eta_full(x, em), f0, scheme=scheme, nsteps=nsteps)
    n = len(em)
    J = np.zeros((n, n))
    for j in range(n):
        h = 1e-6 * max(1.0, abs(em[j]))
        ep = em.copy()
        ep[j] += h
        J[:, j] = (oracle.residual(oracle.eta_full(x, ep), f0, scheme=scheme, nsteps=nsteps)["out"] - F0["out"]) / h
    return J, F0


def test_inverse_jacobian_is_tridiagonal_and_matches_the_model(oracle, fixtures):
    """row-scaled implicit Euler (1D_FEM.c, the benchmarked scheme): inv(J) is tridiagonal and equals the model"""
    N, nsteps = 33, 512
    x = fixtures["n33_x"]
    em = fixtures["n33_eta"][1:-1].copy()
    f0 = oracle.f0_given(x)
    J, F0 = _jacobian(oracle, x, f0, em, oracle.IE_ROWSCALE, nsteps)
    w = np.linalg.eigvalsh(0.5 * (J + J.T))
    assert w.min() > 0 and w.max() < 1.0 and w.max() / w.min() > 100      # smoothing operator: why raw mixing crawls
    H = np.linalg.inv(J)
    n = N - 2
    d0 = np.abs(np.diag(H))
    for k in (2, 3, 5, 8):                                                 # beyond the first off-diagonal: nothing
        assert np.abs(np.diag(H, k)).max() < 2e-3 * d0.min()
    h = x[1] - x[0]
    mid = n // 2
    assert abs(H[mid, mid] * h * h - 1.0) < 0.01 and abs(H[mid, mid + 1] * h * h + 0.5) < 0.01   # (1/2)(-Lap_h) where phi = 1
    # the model, with phi of the same evaluation
    psi = np.sqrt(F0["phi"][1:-1])
    lap = (2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)) / (h * h)
    M = np.eye(n) + 0.5 * lap / np.outer(psi, psi)
    tri = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= 1
    assert np.abs((H - M)[tri]).max() < 0.05 * np.abs(H[tri]).max()        # tridiagonal part: the model to a few percent
    # preconditioned operator: eigenvalues clustered around 1 instead of spread over 2.5 decades
    ev = np.linalg.eigvals(M @ J).real
    assert ev.min() > 0.9 and ev.max() < 3.5


def test_model_clusters_the_spectrum_for_the_consistent_schemes(oracle, fixtures):
    """deal.II matrices (consistent C: implicit Euler and the reference's IRK4): inv(J) carries the inverse mass matrix as well
    (off-diagonals decay by 0.27 per node instead of vanishing), but the same local model still collapses the spectrum of J from
    2.9 decades to less than one, and completed by the inverse mass matrix (J_consistent = J_rowscaled Mass/h) to half a decade"""
    x = fixtures["n33_x"]
    em = fixtures["n33_eta"][1:-1].copy()
    f0 = oracle.f0_given(x)
    n, h = len(em), x[1] - x[0]
    lap = (2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)) / (h * h)
    rho = np.sqrt(3.0) - 2.0      # the 17-point convolution pmix_kernel applies for (Mass/h)^-1 = inverse of tridiag(1,4,1)/6
    conv = np.sqrt(3.0) * sum((rho ** abs(d)) * np.eye(n, k=d) for d in range(-8, 9))
    for scheme in (oracle.IE_CONSISTENT, oracle.IRK4_CONSISTENT):
        J, F0 = _jacobian(oracle, x, f0, em, scheme, 512)
        w = np.linalg.eigvalsh(0.5 * (J + J.T))
        assert w.max() / w.min() > 500
        psi = np.sqrt(F0["phi"][1:-1])
        M = np.eye(n) + 0.5 * lap / np.outer(psi, psi)
        ev = np.linalg.eigvals(M @ J).real
        assert ev.min() > 0.3 and ev.max() < 3.0
        # with the mass-matrix factor the spectrum is the row-scaled scheme's: [1.0, 2.9]
        ev = np.linalg.eigvals(conv @ M @ J).real
        assert ev.min() > 0.95 and ev.max() < 3.2
