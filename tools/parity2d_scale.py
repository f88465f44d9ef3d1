"""y-invariant parity of the 2-D path against the 1-D engine at configuration scale for several CG tolerances.
usage: python tools/parity2d_scale.py [rtol ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scft_b200 as sb
fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
nx, ny, n = 1024, 1023, 2048
e1 = sb.Engine(nx + 1, nsteps=n, scheme=sb.IE_CONSISTENT)
e1.residual(fx["res1024_eta"][1:-1])
phi1, ef = e1.phi(), e1.eta_full()
e1.close()
for rtol in [float(a) for a in sys.argv[1:]] or [1e-12, 1e-13]:
    e2 = sb.Engine2D(nx, ny, nsteps=n, rtol=rtol)
    e2.residual(np.repeat(ef, ny + 1))
    phi2 = e2.phi().reshape(nx + 1, ny + 1)
    it, ms = e2.stats()
    e2.close()
    print(f"rtol {rtol:.0e}: max|phi2D - phi1D| = {np.abs(phi2 - phi1[:, None]).max():.3e}, y-variation {np.abs(phi2 - phi2[:, [0]]).max():.2e}, "
          f"{it} CG iterations = {it / n:.1f}/step, march {ms:.0f} ms")
