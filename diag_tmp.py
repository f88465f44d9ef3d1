import numpy as np, sys, os
sys.path.insert(0,'/root/repo')
import scft_b200
from scft_b200 import sweep
fx=np.load('tests/golden/ref_fixtures.npz')
N=1025
for P in (1,148,296,444,592,888,1332):
    taus,Ls,eta=sweep.make_sweep(0,P,fx['res1024_eta'][1:-1])
    eng=scft_b200.Engine(N,nsteps=2048,scheme=0,max_batch=P)
    eng.set_timing(True)
    for i in range(3): eng.residual(eta)
    eng.march_ms()
    for i in range(5): eng.residual(eta)
    tot,cnt=eng.march_ms()
    ms=tot/cnt
    waves=-(-P//444)
    print("P",P,"ms %.3f"%ms,"cycles/step/wave %.0f"%(ms*1e-3*1.965e9/2048/waves), "DOFsteps/s %.3e"%(P*1023*2048/(ms*1e-3)))
    eng.close()
