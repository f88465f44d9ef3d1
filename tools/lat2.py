"""march-kernel time per launch at P problems (m=1024, n=2048, IE_ROWSCALE): one launch each, CUDA events."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200
from scft_b200 import sweep
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/ref_fixtures.npz'))
N = 1025
Ps = [int(a) for a in sys.argv[1:]] or [1, 148, 296, 444, 592, 1184]
for P in Ps:
    taus, Ls, eta = sweep.make_sweep(0, P, fx['res1024_eta'][1:-1])
    eng = scft_b200.Engine(N, nsteps=2048, scheme=0, max_batch=P)
    for p in range(P):
        eng.set_problem(p, taus[p], Ls[p])
    eng.set_timing(True)
    for i in range(2): eng.residual(eta)
    eng.march_ms()
    for i in range(5): eng.residual(eta)
    tot, cnt = eng.march_ms()
    ms = tot / 5
    print("P %5d slots %d  ms %8.3f  cycles/step %6.0f  DOF-steps/s %.3e  HBM frac %.3f" % (
        P, eng.slots(), ms, ms * 1e-3 * 1.965e9 / 2048, P * 1023 * 2048 / (ms * 1e-3), P * 1023 * 2048 * 8 / (ms * 1e-3) / 6553.9e9), flush=True)
    eng.close()
