"""Ten passes of the 4096-problem sweep on one solver object with SCFTB_SWEEP_TRACE=1 (per-level phase times on stderr):
locates host-side stalls.  usage: SCFTB_SWEEP_TRACE=1 python tools/sweep_trace.py [passes]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scft_b200 import sweep, engine as E
fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
eta33 = fx["n33_eta"][1:-1]
solver = E.SweepSolver(4096, N0=33, levels=6)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    t0 = time.perf_counter()
    r = sweep.converge_block_batched(0, 4096, eta33, levels=6, solver=solver)
    print(f"pass {rep}: {time.perf_counter() - t0:.3f} s (solve {r['seconds']:.3f}, start fields {r['seconds_make_sweep']:.3f})", file=sys.stderr, flush=True)
solver.close()
