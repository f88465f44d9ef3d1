"""The batched continuation with the reference driver's own stepper (IRK4, scft.cc:671-693) to m = 1024:
usage: python tools/sweep_irk4.py [problems=64] [levels=6]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scft_b200 import sweep, engine as E
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
eta33 = fx["n33_eta"][1:-1]
for scheme, name in ((E.IRK4_CONSISTENT, "IRK4 (deal.II driver)"), (E.IE_CONSISTENT, "IE on the deal.II matrices")):
    solver = E.SweepSolver(P, N0=33, levels=levels, scheme=scheme)
    for rep in range(2):
        r = sweep.converge_block_batched(0, P, eta33, levels=levels, solver=solver, scheme=scheme)
    solver.close()
    rows = r["rows"]
    ok = rows[:, 0] == 0
    print(f"{name}: {P} sweep problems to N={(33 - 1) * 2 ** (levels - 1) + 1}: {int(ok.sum())} converged, worst residual {np.nanmax(rows[ok, 1]) if ok.any() else float('nan'):.2e}, "
          f"{r['seconds']:.3f} s; evaluations mean {rows[:, 2].mean():.1f} max {rows[:, 2].max():.0f}, target mesh mean {rows[:, 5].mean():.1f}; "
          f"levels reached {sorted(set(rows[:, 6].astype(int).tolist()))}; F {np.nanmin(rows[:, 4]):.6e} .. {np.nanmax(rows[:, 4]):.6e}")
    for i in np.flatnonzero(~ok)[:5]:
        print(f"   not converged: problem {i} status {int(rows[i, 0])} err {rows[i, 1]:.2e} reached N={int(rows[i, 6])}")
