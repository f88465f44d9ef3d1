"""Parity of the configuration bench.py measures (BASELINE.json configs[2]): the 4096-problem (tau, L, seed) sweep,
N = 1025 nodes, n = 2048 implicit-Euler steps, row-scaled C (1D_FEM.c:95-186), lean (half) history kept per resident
CTA slot, so problems beyond the first wave REUSE history slots.  Checked against the CPU oracle on the problems at
the wave and chunk boundaries of both code paths:
  * scftb_residual_batch with host buffers (>= 4 waves: the chunked, copy-overlapped path bench.py's e2e leg times)
  * scftb_mixer_iterate_device (the device-resident step bench.py's `value` times): three SCFT iterations, the
    march checked at every iterate, the Anderson iterate against the oracle's adm_chen.
Tolerances: BASELINE.json — relative error <= 1e-10 on phi(x) and Q."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, NSTEPS, SCHEME = 1025, 2048, 0
REL = 1e-10


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def check_points(nprob, slots):
    """first/last problem of every wave of resident CTAs plus a few in between"""
    pts = {0, 1, nprob - 1, nprob // 2, nprob // 3}
    for w in range(1, (nprob + slots - 1) // slots):
        pts |= {w * slots - 1, w * slots}
    return sorted(p for p in pts if 0 <= p < nprob)


def oracle_eval(oracle, tau, L, em):
    x = oracle.mesh_uniform(N, L)
    return oracle.residual(oracle.eta_full(x, em), oracle.f0_given(x, tau), scheme=SCHEME, nsteps=NSTEPS, L=L)


def assert_matches(oracle, eng, out_p, p, tau, L, em, tag):
    ref = oracle_eval(oracle, tau, L, em)
    phi = eng.phi(p)
    scale = np.abs(ref["phi"]).max()
    e_phi = np.abs(phi - ref["phi"]).max() / scale
    e_Q = abs(eng.Q(p) - ref["Q"]) / abs(ref["Q"])
    e_out = np.abs(out_p - ref["out"]).max() / scale
    assert e_phi < REL and e_Q < REL and e_out < REL, (tag, p, e_phi, e_Q, e_out)
    return max(e_phi, e_Q, e_out)


def test_sweep_4096_lean_history_host_batch(sb, oracle, fixtures):
    """the exact bench workload through scftb_residual_batch: 4096 problems = 9.2 waves of 444 slots, chunked path"""
    from scft_b200 import sweep
    P = 4096
    taus, Ls, eta = sweep.make_sweep(0, P, fixtures["res1024_eta"][1:-1])
    eng = sb.Engine(N, nsteps=NSTEPS, scheme=SCHEME, max_batch=P)
    for p in range(P):
        eng.set_problem(p, taus[p], Ls[p])
    out = eng.residual(eta)
    assert np.isfinite(out).all()
    slots = eng.slots()
    assert P > 4 * slots, "the chunked path needs at least 4 waves"
    worst = 0.0
    for p in check_points(P, slots):
        worst = max(worst, assert_matches(oracle, eng, out[p], p, taus[p], Ls[p], eta[p], "host_batch"))
    # every problem of a (tau, L) cell shares phi0; a cheap whole-batch property: out + phi = phi0 on every problem
    for p in range(0, P, 97):
        assert np.abs(out[p] + eng.phi(p)[1:-1] - eng.f0_given(p)[1:-1]).max() < 1e-15
    print(f"bench-config parity (host batch): max rel err {worst:.2e} over {len(check_points(P, slots))} problems")
    eng.close()


def test_sweep_lean_history_three_mixer_iterations(sb, oracle, fixtures):
    """>= 1000 problems through the device-resident SCFT iteration (march + Anderson), slots reused across waves"""
    import torch
    from scft_b200 import sweep
    P = 1400
    taus, Ls, eta = sweep.make_sweep(0, P, fixtures["res1024_eta"][1:-1])
    eng = sb.Engine(N, nsteps=NSTEPS, scheme=SCHEME, max_batch=P)
    for p in range(P):
        eng.set_problem(p, taus[p], Ls[p])
    slots = eng.slots()
    pts = check_points(P, slots)
    d_eta = torch.from_numpy(eta).cuda()
    st = torch.cuda.current_stream().cuda_stream
    mixer = sb.AndersonBatch(eng, P, tol=1e-30, lmd=0.99, nn=2)
    mixer.set_freeze(False)
    mixer.reset_device(d_eta.data_ptr(), st)
    worst = 0.0
    for k in range(3):
        xk = mixer.x(st)                 # X_k, the field the next evaluation sees
        if k == 0:
            assert np.array_equal(xk, eta)
        mixer.iterate_device(st)         # Y_k = F(X_k) for all problems, then X_{k+1}
        torch.cuda.synchronize()
        yk = mixer.y(st, k)
        for p in pts:
            worst = max(worst, assert_matches(oracle, eng, yk[p], p, taus[p], Ls[p], xk[p], f"mixer_k{k}"))
    # the iterate after three updates against the oracle's own adm_chen (ADM_chen_C.c) on the oracle residual:
    # before the least-squares amplification sets in (tests/test_gpu_solvers.py) the two paths agree closely
    x3 = mixer.x(st)
    for p in pts[:4]:
        F = lambda em, p=p: oracle_eval(oracle, taus[p], Ls[p], em)["out"]
        _, x_o, _, _ = oracle.adm_chen(F, eta[p], 1e-30, 2, 0.99, 2)
        assert np.abs(x3[p] - x_o).max() < 1e-7 * np.abs(x_o).max(), (p, np.abs(x3[p] - x_o).max())
    print(f"bench-config parity (mixer path): max rel err {worst:.2e} over {len(pts)} problems x 3 iterations")
    mixer.close()
    eng.close()
