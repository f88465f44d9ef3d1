"""Preconditioned Anderson mixing (scftb_pmixer_*, scftb_padm_batch) and the batched continuation (scftb_sweep_*):
converged fields are checked by the CPU oracle (residual, Q, free energy) and against the reference-shaped Broyden flow."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _oracle_check(N, tau, L, eta_mid, scheme=O.IE_ROWSCALE, nsteps=2048):
    x = O.mesh_uniform(N, L)
    ef = O.eta_full(x, eta_mid)
    ref = O.residual(ef, O.f0_given(x, tau), scheme=scheme, nsteps=nsteps, L=L)
    F = O.free_energy(x, ef, tau=tau, L=L, f0bar_=O.f0bar(tau, L))
    return np.max(np.abs(ref["out"])), ref["Q"], F


def test_padm_batch_converges_in_few_evaluations(fixtures):
    """N=129: 16 sweep problems from the interpolated reference field converge in <= 40 evaluations each (the reference's
    staged adm_chen needs thousands, profiles/r1_continuation_m1024.txt) and the oracle confirms every field"""
    import scft_b200 as S
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    N, nprob = 129, 16
    eng = S.Engine(N, nsteps=2048, scheme=S.IE_ROWSCALE, max_batch=nprob)
    x0 = np.zeros((nprob, N - 2))
    pars = []
    for i in range(nprob):
        tau, L, _ = sweep.sweep_params(i * 17)
        pars.append((tau, L))
        eng.set_problem(i, tau, L)
        xs, e = np.linspace(0, L, 33), eta33
        for _ in range(2):
            xs, e = S.refine_mesh(xs, e)
        x0[i] = e
    rc, x, iters, err = S.padm_batch(eng, x0, tol=1e-9, max_iteration=100, nn=10)
    assert rc == 0 and np.all(err < 1e-9) and iters.max() <= 40, (rc, iters, err)
    for i in (0, 5, 15):
        tau, L = pars[i]
        r, Q, _ = _oracle_check(N, tau, L, x[i])
        assert r < 2e-9
        assert abs(Q - eng.Q(i)) <= 1e-10 * abs(Q)
    eng.close()


def test_padm_matches_broyden_fixed_point(fixtures):
    """same discrete problem, two solvers: the preconditioned mixer and the reference-shaped device Broyden agree on the
    converged state (phi, Q to 1e-9; free energy to 1e-9 relative)"""
    import scft_b200 as S
    eta33 = fixtures["n33_eta"][1:-1]
    N = 65
    xs, e = S.refine_mesh(np.linspace(0, S.L_REF, 33), eta33)
    eng = S.Engine(N, nsteps=2048, scheme=S.IE_ROWSCALE, max_batch=N - 2)
    rc, xa, _, erra = S.padm_batch(eng, e[None, :], tol=1e-12, max_iteration=100)
    assert rc == 0 and erra[0] < 1e-12
    eng.residual(xa[0])
    phia, Qa, Fa = eng.phi(0), eng.Q(0), eng.free_energy(0)
    _, check, xb, errb, _ = eng.broydn_device(e, 1e-10, keep_trial=True)
    assert check == 0 and errb < 1e-10
    eng.residual(xb)
    phib, Qb, Fb = eng.phi(0), eng.Q(0), eng.free_energy(0)
    got = (np.max(np.abs(phia - phib)), abs(Qa - Qb) / Qb, abs(Fa - Fb) / abs(Fb), np.max(np.abs(xa[0] - xb)[8:-8]))
    # the field itself is determined up to tol x |J^-1| (6e3 at the wall nodes of this mesh): compare where J is O(1)
    assert got[0] < 1e-9 and got[1] < 1e-9 and got[2] < 1e-9 and got[3] < 1e-6, got
    eng.close()


@pytest.mark.parametrize("scheme_name", ["IE_ROWSCALE", "IE_CONSISTENT", "IRK4_CONSISTENT"])
def test_padm_all_schemes(fixtures, scheme_name):
    import scft_b200 as S
    scheme = getattr(S, scheme_name)
    eta33 = fixtures["n33_eta"][1:-1]
    eng = S.Engine(33, nsteps=2048, scheme=scheme, max_batch=4)
    x0 = np.stack([eta33 * (1 + 0.05 * np.random.default_rng(s).standard_normal(31)) for s in range(4)])
    rc, x, iters, err = S.padm_batch(eng, x0, tol=1e-9, max_iteration=200)
    assert rc == 0 and np.all(err < 1e-9), (rc, iters, err)
    r, _, _ = _oracle_check(33, S.TAU_REF, S.L_REF, x[2], scheme=getattr(O, scheme_name))
    assert r < 2e-9
    eng.close()


def test_refine_batch_equals_host_refine(fixtures):
    import torch
    import scft_b200 as S
    import ctypes as C
    eta33 = fixtures["n33_eta"][1:-1]
    Ls = np.array([3.2, 3.72374357332160, 4.2])
    eta = np.stack([eta33 * (1 + 0.1 * k) for k in range(3)])
    d_L = torch.from_numpy(Ls).cuda()
    d_eta = torch.from_numpy(eta).cuda()
    d_new = torch.zeros((3, 63), dtype=torch.float64, device="cuda")
    rc = S.lib().scftb_refine_uniform_batch_device(3, 33, C.c_void_p(d_L.data_ptr()), C.c_void_p(d_eta.data_ptr()),
                                                   C.c_void_p(d_new.data_ptr()), None)
    assert rc == 0
    got = d_new.cpu().numpy()
    for k in range(3):
        _, ref = S.refine_mesh(np.linspace(0, Ls[k], 33), eta[k])
        assert np.max(np.abs(got[k] - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_free_energy_weights_equal_free_energy(fixtures):
    import scft_b200 as S
    eta33 = fixtures["n33_eta"][1:-1]
    eng = S.Engine(33, nsteps=2048, scheme=S.IRK4_CONSISTENT)
    eng.residual(eta33)
    F = eng.free_energy(0, f0bar=0.0)
    c, f0bar = S.free_energy_weights(33, S.TAU_REF, S.L_REF)
    F2 = (c @ eng.eta_full(0) / f0bar / S.L_REF + np.log(f0bar)) / -1000.0
    assert abs(F - F2) <= 1e-13 * abs(F)
    eng.close()


def test_sweep_solver_continuation_to_257(fixtures):
    """batched continuation 33 -> 65 -> 129 -> 257 of 64 sweep problems: all converge; oracle-checked residual, Q, F; the
    16 seeds of one (tau, L) cell end in the same state"""
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    r = sweep.converge_block_batched(0, 64, eta33, levels=4, want_fields=True)
    rows = r["rows"]
    assert np.all(rows[:, 0] == 0) and np.all(rows[:, 1] < 1e-9) and np.all(rows[:, 6] == 257), rows[rows[:, 0] != 0]
    assert rows[:, 5].max() <= 20            # evaluations on the target mesh
    for p in (0, 17, 63):
        tau, L, _ = sweep.sweep_params(p)
        res, Q, F = _oracle_check(257, tau, L, r["eta"][p])
        assert res < 2e-9
        assert abs(Q - rows[p, 3]) <= 1e-10 * abs(Q)
        assert abs(F - rows[p, 4]) <= 1e-9 * abs(F)
    r2 = sweep.converge_block_batched(256, 258, eta33, levels=4)     # same cells as problems 0, 1; other seeds
    assert np.max(np.abs(r2["rows"][:, 4] - rows[:2, 4])) <= 1e-9 * np.abs(rows[:2, 4]).max()


def test_sweep_solver_to_m1024(fixtures):
    """BASELINE.json configs[2] at full size for a block of the sweep: 48 problems (three (tau, L) rows of the grid) through all
    six levels to N = 1025, n = 2048; every problem converges in <= 12 evaluations on the target mesh; two fields are
    re-evaluated by the CPU oracle"""
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    r = sweep.converge_block_batched(100, 148, eta33, levels=6, want_fields=True)
    rows = r["rows"]
    assert np.all(rows[:, 0] == 0) and np.all(rows[:, 1] < 1e-9) and np.all(rows[:, 6] == 1025), rows[rows[:, 0] != 0]
    assert rows[:, 5].max() <= 12
    for i in (0, 47):
        tau, L, _ = sweep.sweep_params(100 + i)
        res, Q, F = _oracle_check(1025, tau, L, r["eta"][i])
        assert res < 2e-9
        assert abs(Q - rows[i, 3]) <= 1e-10 * abs(Q)
        assert abs(F - rows[i, 4]) <= 1e-9 * abs(F)


def test_padm_edge_cases(fixtures):
    """iteration limit -> SCFTB_ERR_NOCONV with the best iterate returned; NaN start field -> SCFTB_ERR_NAN and the other problems of
    the batch unaffected; window 0 (plain preconditioned relaxation) still descends; non-uniform mesh (Matlab prototype's 59-node
    adaptive mesh) converges and the oracle confirms the field"""
    import scft_b200 as S
    eta33 = fixtures["n33_eta"][1:-1]
    eng = S.Engine(33, nsteps=256, scheme=S.IE_ROWSCALE, max_batch=3)
    x0 = np.stack([eta33 * 1.05, eta33 * 0.95, eta33])
    rc, x, iters, err = S.padm_batch(eng, x0, tol=1e-13, max_iteration=3)       # cannot converge in 4 evaluations
    assert rc == 4 and np.all(np.isfinite(x)) and np.all(err > 1e-13)
    res = np.abs(eng.residual(x)).max(axis=1)
    assert np.all(res <= np.abs(eng.residual(x0)).max(axis=1))                  # the best iterate, not the last one
    bad = x0.copy()
    bad[1, 5] = np.nan
    rc, x, iters, err = S.padm_batch(eng, bad, tol=1e-9, max_iteration=100)
    assert rc == 3 and np.isnan(err[1]) and err[0] < 1e-9 and err[2] < 1e-9
    rc, x, iters, err = S.padm_batch(eng, x0, tol=1e-30, max_iteration=5, nn=0)
    assert rc == 4 and np.all(err <= np.abs(eng.residual(x0)).max(axis=1))
    assert np.allclose(err, np.abs(eng.residual(x)).max(axis=1), rtol=1e-9)   # err is the norm of the returned (best) iterate
    eng.close()
    xs = fixtures["matlab59_x"].copy()
    xs[-1] = max(xs[-1], xs[-2] + 1e-3)
    em = fixtures["matlab59_eta"][1:-1]
    eng = S.Engine(59, nsteps=256, scheme=S.IE_ROWSCALE, tau=0.5302, L=xs[-1], x=xs)
    rc, x, iters, err = S.padm_batch(eng, em[None, :], tol=1e-9, max_iteration=300)
    assert rc == 0 and err[0] < 1e-9, (rc, iters, err)
    ref = O.residual(O.eta_full(xs, x[0]), O.f0_given(xs, 0.5302), scheme=O.IE_ROWSCALE, nsteps=256, L=xs[-1], x=xs)
    assert np.abs(ref["out"]).max() < 2e-9
    eng.close()


def test_sweep_solver_reports_failures_without_touching_the_rest(fixtures):
    """a problem whose start field is NaN is reported (status 2, N reached = 33) and the other problems of the batch converge"""
    import scft_b200 as S
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    taus, Ls, eta0 = sweep.make_sweep(0, 8, eta33)
    eta0[3, :] = np.nan
    solver = S.SweepSolver(8, N0=33, levels=3)
    r = solver.solve(taus, Ls, eta0)
    solver.close()
    rows = r["rows"]
    assert rows[3, 0] == 2 and rows[3, 6] == 33 and np.isnan(rows[3, 3])
    ok = np.delete(np.arange(8), 3)
    assert np.all(rows[ok, 0] == 0) and np.all(rows[ok, 1] < 1e-9) and np.all(rows[ok, 6] == 129)


def test_sweep_solver_consistent_scheme(fixtures):
    """IE on the deal.II matrices (consistent C): the preconditioner carries the inverse mass matrix as well; 32 problems through
    four levels, every problem within 20 evaluations on the target mesh, the oracle confirms a field"""
    import scft_b200 as S
    from scft_b200 import sweep
    eta33 = fixtures["n33_eta"][1:-1]
    r = sweep.converge_block_batched(0, 32, eta33, levels=4, scheme=S.IE_CONSISTENT, want_fields=True)
    rows = r["rows"]
    assert np.all(rows[:, 0] == 0) and np.all(rows[:, 1] < 1e-9) and rows[:, 5].max() <= 20, rows[:, [0, 1, 5]]
    tau, L, _ = sweep.sweep_params(7)
    res, Q, F = _oracle_check(257, tau, L, r["eta"][7], scheme=O.IE_CONSISTENT)
    assert res < 2e-9 and abs(Q - rows[7, 3]) <= 1e-10 * abs(Q) and abs(F - rows[7, 4]) <= 1e-9 * abs(F)
