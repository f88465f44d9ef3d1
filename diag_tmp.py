import numpy as np, sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import scft_b200 as sb
from oracle import oracle as O
from test_gpu_broyden_device import host_broydn
fx=np.load('tests/golden/ref_fixtures.npz')
for (N,scheme,nsteps,scale) in [(33,1,256,1.02),(33,2,256,1.01),(65,0,128,1.0),(129,1,64,1.0),(33,1,128,1.01)]:
    x=O.mesh_uniform(33); em=fx['n33_eta'][1:-1]*scale; Nc=33
    while Nc<N: x,em=sb.refine_mesh(x,em); Nc=2*Nc-1
    eng=sb.Engine(N,nsteps=nsteps,scheme=scheme,max_batch=N-2)
    l0=sb.launch_count()
    h=host_broydn(sb,eng,em,1e-10); l1=sb.launch_count()
    d=eng.broydn_device(em,1e-10); l2=sb.launch_count()
    print(N,scheme,"host rc,chk,err,jc",h[0],h[1],"%.2e"%h[3],h[4],"launches",l1-l0,"| dev",d[0],d[1],"%.2e"%d[3],d[4],"launches",l2-l1,
          "| xdiff %.2e"%np.abs(h[2]-d[2]).max(), "res host %.2e dev %.2e"%(np.abs(eng.residual(h[2])).max(), np.abs(eng.residual(d[2])).max()))
    d2=eng.broydn_device(d[2]+1e-4,1e-10,jc=1)
    print("   reuse:",d2[0],d2[1],"%.2e"%d2[3],d2[4], "xdiff %.2e"%np.abs(d2[2]-d[2]).max())
    eng.close()
