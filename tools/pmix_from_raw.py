"""Preconditioned mixer straight on the finest mesh (no continuation) from the bench's perturbed spectral guess:
residual distribution after k iterations with tol = 0 (no problem ever frozen)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scft_b200 as S
from scft_b200 import sweep
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
taus, Ls, eta = sweep.make_sweep(0, P, fx["res1024_eta"][1:-1])
eng = S.Engine(1025, nsteps=2048, scheme=S.IE_ROWSCALE, max_batch=P)
for p in range(P):
    eng.set_problem(p, taus[p], Ls[p])
m = S.PrecondAndersonBatch(eng, P, tol=0.0, nn=10)
m.reset(eta)
for k in range(60):
    m.iterate_device(0)
    if k in (0, 4, 9, 14, 19, 22, 29, 39, 59):
        done, iters, err = m.status(0)
        print(f"k={k}: finite {int(np.isfinite(err).sum())} median {np.nanmedian(err):.2e} max {np.nanmax(err):.2e} below 1e-9: {int((err < 1e-9).sum())}")
