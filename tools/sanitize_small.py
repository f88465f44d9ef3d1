"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck): pmix_kernel, refine_uniform_kernel,
the wide-chunk march shapes, the matrix-free 2-D rows, the device adm mixer.
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scft_b200 as S
from scft_b200 import sweep
fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
eta33 = fx["n33_eta"][1:-1]
# batched continuation 33 -> 65 -> 129 (pmix_kernel with growing windows, refine kernel, march shapes <1,32>, <2,32>, <4,32>)
r = sweep.converge_block_batched(0, 6, eta33, levels=3, nsteps=64)
print("sweep rows", r["rows"][:, :3].tolist())
# device adm_chen / adm mixers
eng = S.Engine(33, nsteps=64, scheme=1, max_batch=2)
print("adm_chen", eng.adm_chen_batch(np.stack([eta33, eta33 * 1.01]), 1e-30, 12, 0.9, 5)[2].tolist())
print("adm", eng.adm_batch(np.stack([eta33, eta33 * 1.01]), 8)[2].tolist())
eng.close()
# 2-D matrix-free rows, single GPU persistent kernel
e2 = S.Engine2D(16, 5, nsteps=16, rtol=1e-10)
out = e2.residual(np.random.default_rng(0).standard_normal(17 * 6))
print("2d", float(np.abs(out).max()), e2.stats()[0])
e2.close()
