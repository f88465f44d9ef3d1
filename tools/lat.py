"""per-step latency of the march kernel at several occupancies (m=1024, n=2048); usage: lat.py [scheme 0|1|2]"""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200
SCHEME = int(sys.argv[1]) if len(sys.argv) > 1 else 0
from scft_b200 import sweep
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/ref_fixtures.npz'))
N = 1025
for P in (1, 148, 444, 1332, 4096):
    taus, Ls, eta = sweep.make_sweep(0, P, fx['res1024_eta'][1:-1])
    eng = scft_b200.Engine(N, nsteps=2048, scheme=SCHEME, max_batch=P)
    eng.set_timing(True)
    for i in range(2): eng.residual(eta)
    eng.march_ms()
    for i in range(5): eng.residual(eta)
    tot, cnt = eng.march_ms()
    ms = tot / cnt
    waves = -(-P // 444)
    print("P %5d  ms %8.3f  cycles/step/wave %6.0f  DOF-steps/s %.3e" % (P, ms, ms * 1e-3 * 1.965e9 / 2048 / waves, P * 1023 * 2048 / (ms * 1e-3)))
    eng.close()
