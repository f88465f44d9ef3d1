"""How many problems of the 4096-problem sweep survive the first iterations of a device-resident Anderson schedule
(lmd = relaxation, nn = window)?  Counts finite residuals and the error quantiles after `its` iterations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scft_b200  # noqa: E402
from scft_b200 import sweep  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
its = int(sys.argv[2]) if len(sys.argv) > 2 else 25
fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))
taus, Ls, eta = sweep.make_sweep(0, P, fx["res1024_eta"][1:-1])
eng = scft_b200.Engine(1025, nsteps=2048, scheme=0, max_batch=P)
for p in range(P):
    eng.set_problem(p, taus[p], Ls[p])
d_eta = torch.from_numpy(eta).cuda()
st = torch.cuda.current_stream().cuda_stream
for nn, lmd in [(2, 0.99), (1, 0.99), (0, 0.99), (2, 0.9), (3, 0.9), (1, 0.9), (0, 0.9), (2, 0.999)]:
    mixer = scft_b200.AndersonBatch(eng, P, tol=1e-30, lmd=lmd, nn=nn)
    mixer.set_freeze(False)
    mixer.reset_device(d_eta.data_ptr(), st)
    errs = []
    for k in range(its):
        mixer.iterate_device(st)
        if k in (0, 4, 9, 14, 19, its - 1):
            done, iters, err = mixer.status(st)
            errs.append((k, int(np.isfinite(err).sum()), float(np.nanmedian(err)), float(np.nanmax(err))))
    print(f"nn={nn} lmd={lmd}: " + " | ".join(f"k={k}: finite {f} med {m:.2e} max {x:.2e}" for k, f, m, x in errs), flush=True)
    mixer.close()
