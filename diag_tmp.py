import sys, os, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O
from scft_b200 import sweep, engine as E
fx = dict(np.load("tests/golden/ref_fixtures.npz"))
eta33 = fx["n33_eta"][1:-1]

def cmp(tag, N, tau, L, eta, scheme, n, nonuni=False):
    x = O.mesh_uniform(N, L)
    eng = E.Engine(N, nsteps=n, scheme=scheme, tau=tau, L=L, x=(x if nonuni else None))
    out = eng.residual(eta); phi = eng.phi(0)
    ref = O.residual(O.eta_full(x, eta), O.f0_given(x, tau), scheme=scheme, nsteps=n, L=L, x=(x if nonuni else None))
    print(f"{tag}: N={N} scheme={scheme} n={n} nonuni={nonuni} max|out_gpu|={np.abs(out).max():.3e} max|out_orc|={np.abs(ref['out']).max():.3e} "
          f"max|dout|={np.abs(out-ref['out']).max():.3e} max|dphi|={np.abs(phi-ref['phi']).max():.3e} dQ={abs(eng.Q(0)-ref['Q']):.2e} "
          f"df0={np.abs(eng.f0_given(0)-O.f0_given(x,tau)).max():.2e}", flush=True)
    eng.close()

for p in (0, 2):
    tau, L, seed = sweep.sweep_params(p)
    z = np.random.default_rng(seed).standard_normal(31)
    cont = sweep.Continuation(N_target=129, tol=1e-9)
    r = cont.solve(tau, L, eta33 * (1 + 0.05 * z)); cont.close()
    print("p", p, "tau", tau, "L", L, "cont129: check", r["check"], "err", r["err"], "levels", r["level_err"], flush=True)
    eta = r["eta_mid"]
    cmp("conv", 129, tau, L, eta, 0, 2048)
    cmp("conv", 129, tau, L, eta, 0, 2048, nonuni=True)
    cmp("conv", 129, tau, L, eta, 1, 2048)
    cmp("conv", 129, tau, L, eta, 0, 64)
    cmp("conv-reftauL", 129, E.TAU_REF, E.L_REF, eta, 0, 2048)
    cmp("rand", 129, tau, L, np.random.default_rng(1).standard_normal(127) * 2, 0, 2048)

# the m=1024 stall
for p in (0, 1, 2):
    tau, L, seed = sweep.sweep_params(p)
    z = np.random.default_rng(seed).standard_normal(31)
    cont = sweep.Continuation(N_target=1025, tol=1e-9)
    t0 = time.perf_counter()
    r = cont.solve(tau, L, eta33 * (1 + 0.05 * z))
    print("p", p, "tau", tau, "L", L, "cont1025: check", r["check"], "err", r["err"], "N", r["N"], "levels", ["%.2e" % e for e in r["level_err"]],
          "%.2fs" % (time.perf_counter() - t0), flush=True)
    N = r["N"]
    cmp("stalled" if r["check"] else "conv", N, tau, L, r["eta_mid"], 0, 2048)
    if r["check"]:
        # host-flow Broyden (host QR) on the same engine from the stalled field
        eng = cont.engines[cont.levels.index(N)]
        eng.bind_global()
        Lb = E.lib()
        x = r["eta_mid"].copy()
        for attempt in range(2):
            chk, err, jc = C.c_int(1), C.c_double(1e-9), C.c_int(0)
            t0 = time.perf_counter()
            rc = Lb.scftb_broydn(Lb.scftb_callback_c0, x.ctypes.data_as(C.POINTER(C.c_double)), N - 2, C.byref(chk), C.byref(err), C.byref(jc))
            print("   host-flow broydn from the stalled field: rc", rc, "check", chk.value, "err", err.value, "%.2fs" % (time.perf_counter() - t0), flush=True)
            if chk.value == 0: break
    cont.close()
