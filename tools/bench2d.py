"""2-D path throughput: DOF-steps/s, CG iterations per step, time per iteration and bytes/DOF/iteration.
usage: python tools/bench2d.py [nx ny nsteps]   (single GPU)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scft_b200 as sb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1023
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1023
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_fixtures.npz"))
L = sb.L_REF
x = L * np.arange(nx + 1) / nx
# eta(x,y) = spline of the m=1024 field in x, modulated in y
xr = fx["res1024_xl"] * L
eta_x = np.interp(x, xr, fx["res1024_eta"])
y = np.arange(ny + 1) / ny
eta = (eta_x[:, None] * (1 + 0.1 * np.cos(2 * np.pi * y)[None, :])).ravel()
eng = sb.Engine2D(nx, ny, L=L, Ly=L, nsteps=n, rtol=1e-12)
for rep in range(2):
    t0 = time.perf_counter(); eng.residual(eta); wall = time.perf_counter() - t0
    it, ms = eng.stats()
ndof = (nx + 1) * (ny + 1)
print(f"mesh {nx}x{ny} ({ndof} DOFs), {n} steps: march {ms:.1f} ms (wall {wall*1e3:.1f}), {it} CG iterations = {it/n:.1f}/step, "
      f"{ms*1e3/it:.2f} us/iteration, {ndof*n/(ms*1e-3):.3e} DOF-steps/s, "
      f"{128*ndof*it/(ms*1e-3)/1e9:.0f} GB/s at 128 B/DOF/iteration (matrix-free rows)")
