"""One launch of the IRK4 march kernel for an ncu capture (m=1024, n=2048, P problems, default 296 = one wave of the
tensor-memory kernel).  usage: ncu --set full --import-source on -k regex:irk4 -c 1 -o OUT python tools/prof_irk4.py [P]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200  # noqa: E402
from scft_b200 import sweep  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_fixtures.npz"))
taus, Ls, eta = sweep.make_sweep(0, P, fx["res1024_eta"][1:-1])
eng = scft_b200.Engine(1025, nsteps=2048, scheme=scft_b200.IRK4_CONSISTENT, max_batch=P)
for p in range(P):
    eng.set_problem(p, taus[p], Ls[p])
out = eng.residual(eta)
print(eng.kernel_name(), P, "finite outputs:", int(np.isfinite(out).all(axis=1).sum()))
eng.close()
