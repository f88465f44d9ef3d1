"""ctypes front-end of the CPU oracle (oracle/scft_oracle.c) and of the reference's own C
compiled into oracle/_ref (oracle/ref_shim.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package scft_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libscft_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libscft_ref.so")
REF_ROOT_SO = os.path.join(HERE, "_ref", "libscft_ref_root.so")

IE_ROWSCALE, IE_CONSISTENT, IRK4_CONSISTENT = 0, 1, 2
QUAD_ROMBERG, QUAD_TRAPEZOID = 0, 1
SPLINE_NATURAL, SPLINE_NOTAKNOT, SPLINE_GIVEN = 0, 1, 2

# the reference's physical parameters (drivescft.cc:269)
TAU_REF = 5.30252230020752e-01
L_REF = 3.72374357332160

_dp = C.POINTER(C.c_double)
FUNC = C.CFUNCTYPE(None, C.c_int, _dp, _dp)


class Config(C.Structure):
    _fields_ = [("scheme", C.c_int), ("N", C.c_int), ("nsteps", C.c_int), ("quadrature", C.c_int),
                ("sign", C.c_double), ("L", C.c_double), ("x", _dp)]


def build(ref=True):
    """(Re)build the oracle and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE] + ([] if ref else [os.path.join(HERE, "_build", "libscft_oracle.so")]),
                   check=True, stdout=subprocess.DEVNULL)


def _p(a):
    return a.ctypes.data_as(_dp)


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_romint.restype = C.c_double
        L.orc_romint.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_romberg_weights.argtypes = [C.c_int, C.c_double, _dp]
        L.orc_f0_given.argtypes = [C.c_int, _dp, C.c_double, _dp]
        L.orc_f0bar.restype = C.c_double
        L.orc_f0bar.argtypes = [C.c_double, C.c_double]
        L.orc_spline.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.orc_eta_full.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_residual.restype = C.c_int
        L.orc_residual.argtypes = [C.POINTER(Config), _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_residual_ab.restype = C.c_int
        L.orc_residual_ab.argtypes = [C.POINTER(Config), _dp, _dp, C.c_int, C.c_double, _dp, _dp, _dp, _dp,
                                      C.POINTER(C.c_double)]
        L.orc_free_energy.restype = C.c_double
        L.orc_free_energy.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_gaussj.restype = C.c_int
        L.orc_gaussj.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int]
        L.orc_adm_chen.restype = C.c_int
        L.orc_adm_chen.argtypes = [FUNC, _dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                   _dp, C.POINTER(C.c_int)]
        L.orc_adm.restype = C.c_int
        L.orc_adm.argtypes = [FUNC, _dp, C.c_int, C.c_int, _dp, C.POINTER(C.c_int)]
        L.orc_fast_sweep.restype = C.c_int
        L.orc_fast_sweep.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp,
                                     C.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------------
def romint(f, hh):
    f = _arr(f)
    return lib().orc_romint(_p(f), len(f) - 1, hh)


def romberg_weights(m, hh):
    w = np.zeros(m + 1)
    lib().orc_romberg_weights(m, hh, _p(w))
    return w


def mesh_uniform(N, L=L_REF):
    return L * np.arange(N) / (N - 1)


def f0_given(x, tau=TAU_REF):
    x = _arr(x)
    out = np.zeros_like(x)
    lib().orc_f0_given(len(x), _p(x), tau, _p(out))
    return out


def f0bar(tau=TAU_REF, L=L_REF):
    return lib().orc_f0bar(tau, L)


def spline(x, y, xp, mode=SPLINE_NATURAL, bc=0.0):
    x, y, xp = _arr(x), _arr(y), _arr(xp)
    yp = np.zeros_like(xp)
    lib().orc_spline(_p(x), _p(y), _p(xp), _p(yp), len(x), len(xp), mode, bc)
    return yp


def eta_full(x, eta_mid):
    x, eta_mid = _arr(x), _arr(eta_mid)
    out = np.zeros(len(x))
    lib().orc_eta_full(len(x), _p(x), _p(eta_mid), _p(out))
    return out


def residual(eta_full_, f0, scheme=IE_CONSISTENT, nsteps=2048, L=L_REF, x=None, quadrature=QUAD_ROMBERG,
             sign=1.0, want_hist=False):
    """Returns dict(out=..., phi=..., Q=..., hist=... (N x (n+1)) if want_hist)."""
    eta_full_, f0 = _arr(eta_full_), _arr(f0)
    N = len(eta_full_)
    xa = None if x is None else _arr(x)
    cfg = Config(scheme, N, nsteps, quadrature, sign, L, _p(xa) if xa is not None else None)
    out = np.zeros(N - 2)
    phi = np.zeros(N)
    hist = np.zeros((N, nsteps + 1)) if want_hist else None
    Q = C.c_double(0)
    lib().orc_residual(C.byref(cfg), _p(eta_full_), _p(f0), _p(out), _p(phi),
                       _p(hist) if want_hist else None, C.byref(Q))
    r = dict(out=out, phi=phi, Q=Q.value)
    if want_hist:
        r["hist"] = hist
    return r


def fast_sweep(taus, Ls, eta_mid, N, scheme=IE_ROWSCALE, nsteps=2048, quadrature=QUAD_ROMBERG, sign=1.0, threads=1,
               want_phi=True):
    """The "fair CPU" baseline (oracle/scft_fast.c): Thomas once per field, half history, 8 problems interleaved for
    SIMD, `threads` POSIX threads.  eta_mid [nprob, N-2] -> dict(out [nprob, N-2], phi [nprob, N], Q [nprob])."""
    taus, Ls, eta_mid = _arr(taus), _arr(Ls), _arr(eta_mid)
    nprob = len(taus)
    assert eta_mid.shape == (nprob, N - 2)
    out = np.zeros((nprob, N - 2))
    phi = np.zeros((nprob, N)) if want_phi else None
    Q = np.zeros(nprob)
    rc = lib().orc_fast_sweep(nprob, N, nsteps, scheme, quadrature, sign, _p(taus), _p(Ls), _p(eta_mid), _p(out),
                              _p(phi) if want_phi else None, _p(Q), int(threads))
    if rc:
        raise ValueError("orc_fast_sweep: IE schemes on uniform meshes only (Romberg needs nsteps = 2^k >= 16)")
    return dict(out=out, phi=phi, Q=Q)


def residual_ab(etaA_full, etaB_full, jf, chiN, f0, scheme=IE_ROWSCALE, nsteps=2048, L=L_REF, x=None,
                quadrature=QUAD_ROMBERG, sign=1.0):
    """Two-species residual (orc_residual_ab): dict(out[2*(N-2)], phiA, phiB, Q)."""
    a, b, f0 = _arr(etaA_full), _arr(etaB_full), _arr(f0)
    N = len(a)
    xa = None if x is None else _arr(x)
    cfg = Config(scheme, N, nsteps, quadrature, sign, L, _p(xa) if xa is not None else None)
    out, pa, pb, Q = np.zeros(2 * (N - 2)), np.zeros(N), np.zeros(N), C.c_double(0)
    rc = lib().orc_residual_ab(C.byref(cfg), _p(a), _p(b), int(jf), float(chiN), _p(f0), _p(out), _p(pa), _p(pb),
                               C.byref(Q))
    if rc:
        raise ValueError("orc_residual_ab: IE schemes only, 0 < jf < nsteps")
    return dict(out=out, phiA=pa, phiB=pb, Q=Q.value)


def free_energy(x, eta_full_, tau=TAU_REF, L=L_REF, f0bar_=0.892581217773656, nplot=(1 << 18) + 1):
    x, eta_full_ = _arr(x), _arr(eta_full_)
    return lib().orc_free_energy(len(x), _p(x), _p(eta_full_), tau, L, f0bar_, nplot)


def gaussj(a, b, variant=0):
    a = _arr(a).copy()
    b = _arr(b).copy().reshape(a.shape[0], -1)
    rc = lib().orc_gaussj(_p(a), a.shape[0], _p(b), b.shape[1], variant)
    return rc, a, b


def _wrap(pyfunc, n):
    def cb(nn, pin, pout):
        xin = np.ctypeslib.as_array(pin, shape=(nn,))
        res = pyfunc(xin.copy())
        np.ctypeslib.as_array(pout, shape=(nn,))[:] = res
    return FUNC(cb)


def adm_chen(pyfunc, x0, tol, max_iteration, lmd, nn, final=False):
    x = _arr(x0).copy()
    trace = np.full(max_iteration + 2, np.nan)
    it = C.c_int(0)
    rc = lib().orc_adm_chen(_wrap(pyfunc, len(x)), _p(x), tol, max_iteration, len(x), lmd, nn, int(final),
                            _p(trace), C.byref(it))
    return rc, x, trace[: it.value + 1], it.value


def adm(pyfunc, x0, maxits=100000):
    x = _arr(x0).copy()
    trace = np.full(maxits + 1, np.nan)
    it = C.c_int(0)
    rc = lib().orc_adm(_wrap(pyfunc, len(x)), _p(x), len(x), maxits, _p(trace), C.byref(it))
    return rc, x, trace[: it.value], it.value


# ------------------------------------------------------------------------------------------------
# The reference's own C (oracle/_ref).  Present in the build container and shipped prebuilt.
_ref = None
_ref_root = None


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF_ROOT_SO)


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_SO)
        R.ref_romint.restype = C.c_double
        R.ref_romint.argtypes = [_dp, C.c_int, C.c_double]
        R.ref_spline_chen.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double]
        R.ref_gaussj.restype = C.c_int
        R.ref_gaussj.argtypes = [_dp, C.c_int, _dp, C.c_int]
        R.ref_adm_chen.restype = C.c_int
        R.ref_adm_chen.argtypes = [FUNC, _dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        R.ref_broydn.restype = C.c_int
        R.ref_broydn.argtypes = [FUNC, _dp, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]
        R.ref_broydn_raw.restype = C.c_int
        R.ref_broydn_raw.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]
        _ref = R
    return _ref


def ref_root():
    global _ref_root
    if _ref_root is None:
        R = C.CDLL(REF_ROOT_SO)
        R.ref_adm.restype = C.c_int
        R.ref_adm.argtypes = [FUNC, _dp, C.c_int]
        _ref_root = R
    return _ref_root


def ref_romint(f, hh):
    f = _arr(f)
    return ref().ref_romint(_p(f), len(f) - 1, hh)


def ref_spline(x, y, xp, mode=SPLINE_NATURAL, bc=0.0):
    x, y, xp = _arr(x), _arr(y), _arr(xp)
    yp = np.zeros_like(xp)
    ref().ref_spline_chen(_p(x), _p(y), _p(xp), _p(yp), len(x), len(xp), mode, bc)
    return yp


def ref_gaussj(a, b):
    a = _arr(a).copy()
    b = _arr(b).copy().reshape(a.shape[0], -1)
    rc = ref().ref_gaussj(_p(a), a.shape[0], _p(b), b.shape[1])
    return rc, a, b


def ref_adm_chen(pyfunc, x0, tol, max_iteration, lmd, nn, final=False):
    x = _arr(x0).copy()
    rc = ref().ref_adm_chen(_wrap(pyfunc, len(x)), _p(x), tol, max_iteration, len(x), lmd, nn, int(final))
    return rc, x


def ref_broydn(pyfunc, x0, tolf, jc=0):
    x = _arr(x0).copy()
    err = C.c_double(0)
    jcv = C.c_int(jc)
    check = ref().ref_broydn(_wrap(pyfunc, len(x)), _p(x), len(x), tolf, C.byref(err), C.byref(jcv))
    return check, x, err.value, jcv.value


def ref_broydn_raw(fn_addr, x0, tolf, jc=0):
    """the reference's broydn.c driving a RAW C callback (address of a void f(int, double[1..n], double[1..n]),
    e.g. scftb_callback_nr1) — no Python or C adapter between the reference and the callback"""
    x = _arr(x0).copy()
    err = C.c_double(0)
    jcv = C.c_int(jc)
    check = ref().ref_broydn_raw(C.c_void_p(fn_addr), _p(x), len(x), tolf, C.byref(err), C.byref(jcv))
    return check, x, err.value, jcv.value


def ref_adm_chen_raw(fn_addr, x0, tol, max_iteration, lmd, nn, final=False):
    """the reference's adm_chen (ADM_chen_C.c:18) driving a RAW 0-based C callback (e.g. scftb_callback_c0)"""
    x = _arr(x0).copy()
    f = C.cast(C.c_void_p(fn_addr), FUNC)   # a function-pointer VALUE: ctypes passes the address itself
    rc = ref().ref_adm_chen(f, _p(x), tol, max_iteration, len(x), lmd, nn, int(final))
    return rc, x


def ref_adm(pyfunc, x0):
    x = _arr(x0).copy()
    rc = ref_root().ref_adm(_wrap(pyfunc, len(x)), _p(x), len(x))
    return rc, x


# ------------------------------------------------------------------------------------------------
# fixture readers (formats: scft_util.cc:13-41 and 1D_FEM.c:322-342)
def read_yita_file(path):
    """'N= %d, ERROR= %e' / 'mean_field_free_energy, %f' / rows 'i,x,eta' -> dict."""
    with open(path) as fh:
        lines = fh.read().strip().splitlines()
    head = lines[0]
    N = int(head.split("N=")[1].split(",")[0])
    err = float(head.split("ERROR=")[1]) if "ERROR=" in head else None
    start = 1
    F = None
    if "mean_field_free_energy" in lines[1]:
        F = float(lines[1].split(",")[1])
        start = 2
    x = np.zeros(N)
    eta = np.zeros(N)
    for ln in lines[start:]:
        i, xv, v = ln.split(",")
        x[int(i)] = float(xv)
        eta[int(i)] = float(v)
    return dict(N=N, error=err, F=F, x=x, eta=eta)


def read_res_file(path):
    """Q. Wang's spectral .res file: 9 header lines then rows x/l, phi, eta, phie, phij."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    hdr = " ".join(lines[:6])
    meta = {}
    for key in ("l", "mphi", "t", "Z", "f"):
        meta[key] = float(hdr.split(key + "=")[1].split()[0])
    rows = [ln.split() for ln in lines[9:] if ln.strip()]
    a = np.array(rows, dtype=np.float64)
    return dict(meta=meta, xl=a[:, 0], phi=a[:, 1], eta=a[:, 2])
