"""Time-to-converge flow of the sweep (scft_b200.sweep.Continuation): continuation N=33 -> 65 -> 129 with the
device-resident Broyden solver; the converged fields are checked with the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [0, 15, 240, 255, 1000])
def test_continuation_converges_sweep_problem(fixtures, p):
    from scft_b200 import sweep, engine as E
    eta33 = fixtures["n33_eta"][1:-1]
    tau, L, seed = sweep.sweep_params(p)
    z = np.random.default_rng(seed).standard_normal(len(eta33))
    cont = sweep.Continuation(N_target=129, N0=33, nsteps=2048, scheme=E.IE_ROWSCALE, tol=1e-9)
    try:
        r = cont.solve(tau, L, eta33 * (1 + 0.05 * z))
    finally:
        cont.close()
    assert r["check"] == 0 and r["N"] == 129 and r["err"] < 1e-9, r
    x = O.mesh_uniform(129, L)
    ef = O.eta_full(x, r["eta_mid"])
    ref = O.residual(ef, O.f0_given(x, tau), scheme=O.IE_ROWSCALE, nsteps=2048, L=L)
    assert np.max(np.abs(ref["out"])) < 2e-9          # the oracle agrees that this field is converged
    assert abs(ref["Q"] - r["Q"]) <= 1e-10 * abs(ref["Q"])
    F = O.free_energy(x, ef, tau=tau, L=L, f0bar_=O.f0bar(tau, L))
    assert abs(F - r["F"]) <= 1e-9 * abs(F)


def test_continuation_rejects_unreachable_target():
    from scft_b200 import sweep
    with pytest.raises(ValueError):
        sweep.Continuation(N_target=100, N0=33)


def test_threaded_block_equals_serial(fixtures):
    """host threads only overlap independent problems: same rows as the serial loop"""
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    a = sweep.converge_block(3, 9, eta33, N_target=65, threads=1)
    b = sweep.converge_block(3, 9, eta33, N_target=65, threads=3)
    assert np.all(a[:, 0] == 0) and np.array_equal(a[:, :4], b[:, :4])
