"""One launch of the march kernel for an ncu capture (m=1024, n=2048 sweep problems).
usage:  ncu --set full --clock-control none --import-source on -k regex:march_ie -c 1 -o OUT python tools/prof_march.py [one|ab] [P]
one: the one-sweep residual of P (default 4096) problems;  ab: the two-species march of P (default 1332) problems."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200  # noqa: E402
from scft_b200 import sweep  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "one"
P = int(sys.argv[2]) if len(sys.argv) > 2 else (4096 if mode == "one" else 1332)
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_fixtures.npz"))
taus, Ls, eta = sweep.make_sweep(0, P, fx["res1024_eta"][1:-1])
eng = scft_b200.Engine(1025, nsteps=2048, scheme=scft_b200.IE_ROWSCALE, max_batch=P)
for p in range(P):
    eng.set_problem(p, taus[p], Ls[p])
if mode == "ab":
    eng.set_diblock(0.25, 10.0)
    out = eng.residual_ab(np.concatenate([eta, 0.9 * eta], axis=1))
else:
    out = eng.residual(eta)
print(mode, P, "finite outputs:", int(np.isfinite(out).all(axis=1).sum()))
eng.close()
