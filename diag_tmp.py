import numpy as np, sys
sys.path.insert(0,'/root/repo')
import scft_b200 as sb
from oracle import oracle as O
fx=np.load('tests/golden/ref_fixtures.npz')
N=33; x=O.mesh_uniform(N); em=fx['res32_eta'][1:-1]
for n in (33,7,3,4,5,2):
    for scheme in (0,1,2):
        for quad in (1,):
            eng=sb.Engine(N,nsteps=n,scheme=scheme,quadrature=quad)
            eng.residual(em)
            ref=O.residual(O.eta_full(x,em),O.f0_given(x),scheme=scheme,nsteps=n,quadrature=quad)
            print(n,scheme,"rel err %.2e"%(np.abs(eng.phi()-ref['phi']).max()/np.abs(ref['phi']).max()))
            eng.close()
