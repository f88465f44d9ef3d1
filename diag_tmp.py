import numpy as np, sys, os
sys.path.insert(0,'/root/repo')
import torch, scft_b200
from scft_b200 import sweep
fx=np.load('tests/golden/ref_fixtures.npz')
P=1024; N=1025
taus,Ls,eta=sweep.make_sweep(0,P,fx['res1024_eta'][1:-1])
for scheme in (0,1):
  for (lmd,nn) in [(0.9,3),(0.99,2),(0.99,3),(0.9,15)]:
    eng=scft_b200.Engine(N,nsteps=2048,scheme=scheme,max_batch=P)
    for p in range(P): eng.set_problem(p,taus[p],Ls[p])
    mx=scft_b200.AndersonBatch(eng,P,tol=1e-30,lmd=lmd,nn=nn)
    mx.reset(eta)
    hist=[]
    for k in range(30):
        mx.iterate_device(0)
        if k in (0,1,2,5,10,20,29):
            done,iters,err=mx.status(0)
            hist.append((k,int((done!=0).sum()), float(np.nanmedian(err)), float(np.nanmax(err))))
    print("scheme",scheme,"lmd",lmd,"nn",nn,hist)
    mx.close(); eng.close()
