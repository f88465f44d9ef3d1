"""The "fair CPU" baseline (oracle/scft_fast.c: Thomas once per field, half history, SIMD over interleaved problems,
threads over groups) against the checker oracle (orc_residual: pivoting band LU, full history, romint.c per node).
It is what bench.py's cpu_baseline / --impl reference legs time, so it has to compute the same thing."""
import numpy as np
import pytest


@pytest.mark.parametrize("scheme", [0, 1])
@pytest.mark.parametrize("N,nsteps,nprob,threads", [(33, 2048, 3, 1), (65, 64, 19, 4), (129, 32, 8, 2), (1025, 64, 9, 3)])
def test_fast_sweep_matches_oracle(oracle, fixtures, scheme, N, nsteps, nprob, threads):
    rng = np.random.default_rng(N + nsteps + scheme)
    taus = rng.uniform(0.40, 0.66, nprob)
    Ls = rng.uniform(3.2, 4.2, nprob)
    if N == 33:
        base = fixtures["res32_eta"][1:-1]
    elif N == 1025:
        base = fixtures["res1024_eta"][1:-1]
    else:
        base = 3.0 * rng.standard_normal(N - 2)
    eta = base[None, :] * (1 + 0.05 * rng.standard_normal((nprob, N - 2)))
    r = oracle.fast_sweep(taus, Ls, eta, N, scheme=scheme, nsteps=nsteps, threads=threads)
    for p in range(nprob):
        x = oracle.mesh_uniform(N, Ls[p])
        ref = oracle.residual(oracle.eta_full(x, eta[p]), oracle.f0_given(x, taus[p]), scheme=scheme, nsteps=nsteps, L=Ls[p])
        scale = np.abs(ref["phi"]).max()
        assert np.abs(r["phi"][p] - ref["phi"]).max() < 1e-11 * scale
        assert np.abs(r["out"][p] - ref["out"]).max() < 1e-11 * scale
        assert abs(r["Q"][p] - ref["Q"]) < 1e-11 * abs(ref["Q"])


def test_fast_sweep_trapezoid_sign_and_bad_arguments(oracle, fixtures):
    N, n = 33, 100
    eta = fixtures["res32_eta"][1:-1][None, :]
    r = oracle.fast_sweep([oracle.TAU_REF], [oracle.L_REF], eta, N, scheme=0, nsteps=n, quadrature=1, sign=-1.0)
    x = oracle.mesh_uniform(N)
    ref = oracle.residual(oracle.eta_full(x, eta[0]), oracle.f0_given(x), scheme=0, nsteps=n, quadrature=1, sign=-1.0)
    assert np.abs(r["out"][0] - ref["out"]).max() < 1e-12
    with pytest.raises(ValueError):
        oracle.fast_sweep([0.5], [3.7], eta, N, scheme=2, nsteps=64)          # IRK4 is not an IE scheme
    with pytest.raises(ValueError):
        oracle.fast_sweep([0.5], [3.7], eta, N, scheme=0, nsteps=100)         # Romberg needs 2^k steps
