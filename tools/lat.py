"""per-step latency of the march kernel at several occupancies (m=1024, n=2048); usage: lat.py [scheme 0|1|2] [ab]
ab: the two-species march (q and q+ as separate sweeps, P = 2 propagator sweeps per evaluation, fA = 1/4, chiN = 10)"""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200
SCHEME = int(sys.argv[1]) if len(sys.argv) > 1 else 0
from scft_b200 import sweep
AB = len(sys.argv) > 2 and sys.argv[2] == 'ab'
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/ref_fixtures.npz'))
N = 1025
for P in ((1, 148, 444, 1332) if AB else (1, 148, 444, 1332, 4096)):
    taus, Ls, eta = sweep.make_sweep(0, P, fx['res1024_eta'][1:-1])
    eng = scft_b200.Engine(N, nsteps=2048, scheme=SCHEME, max_batch=P)
    eng.set_timing(True)
    if AB:
        eng.set_diblock(0.25, 10.0)
        eta = np.concatenate([eta, 0.9 * eta], axis=1)
    run = eng.residual_ab if AB else eng.residual
    for i in range(2): run(eta)
    eng.march_ms()
    for i in range(5): run(eta)
    tot, cnt = eng.march_ms()
    ms = tot / 5   # per call: large host batches run as several chunk launches (scftb_residual_batch pipeline)
    waves = -(-P // 444)
    sweeps = 2 if AB else 1
    print("P %5d  ms %8.3f  cycles/step/wave %6.0f  DOF-steps/s %.3e%s" % (P, ms, ms * 1e-3 * 1.965e9 / (2048 * sweeps) / waves,
          sweeps * P * 1023 * 2048 / (ms * 1e-3), "  (two-species, 2 sweeps)" if AB else ""))
    eng.close()
