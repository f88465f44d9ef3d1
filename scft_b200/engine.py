"""ctypes host mirror of include/scft_b200.h.

Engine mirrors the role of SCFT::HeatEquation<2> (DEALII_SCFT/include/SCFT.h:117-155): construct
with (tau, N, total_time_step, L) semantics, call run()/residual() with the interior field.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SCFTB_LIB: another build of the same library (kernel-variant experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("SCFTB_LIB") or os.path.join(HERE, "lib", "libscft_b200.so")

IE_ROWSCALE, IE_CONSISTENT, IRK4_CONSISTENT = 0, 1, 2
QUAD_ROMBERG, QUAD_TRAPEZOID = 0, 1
TAU_REF = 5.30252230020752e-01   # drivescft.cc:269
L_REF = 3.72374357332160

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
FUNC = C.CFUNCTYPE(None, C.c_int, _dp, _dp)


class ScftError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("scheme", C.c_int), ("N", C.c_int), ("nsteps", C.c_int), ("quadrature", C.c_int),
                ("sign", C.c_double), ("max_batch", C.c_int), ("device", C.c_int), ("store_history", C.c_int)]


class _SweepConfig(C.Structure):
    _fields_ = [("scheme", C.c_int), ("N0", C.c_int), ("levels", C.c_int), ("nsteps", C.c_int), ("quadrature", C.c_int),
                ("tol", C.c_double), ("nn", C.c_int), ("maxit", C.c_int), ("cap", C.c_double), ("device", C.c_int)]


SWEEP_COLS = 7


class _Config2D(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("L", C.c_double), ("Ly", C.c_double), ("tau", C.c_double),
                ("nsteps", C.c_int), ("quadrature", C.c_int), ("sign", C.c_double), ("rtol", C.c_double),
                ("maxit", C.c_int), ("device", C.c_int), ("rank", C.c_int), ("world", C.c_int),
                ("store_history", C.c_int)]


_lib = None

# every symbol include/scft_b200.h declares
EXPORTS = ["scftb_create", "scftb_destroy", "scftb_last_error", "scftb_launch_count", "scftb_set_problem",
           "scftb_residual", "scftb_residual_batch", "scftb_residual_batch_device", "scftb_get_phi", "scftb_get_Q",
           "scftb_get_f0_given", "scftb_get_eta_full", "scftb_get_q_history", "scftb_free_energy",
           "scftb_bind_global", "scftb_callback_nr1", "scftb_callback_c0", "scftb_callback_fixedpoint_c0",
           "scftb_funcerr", "scftb_adm_chen", "scftb_adm", "scftb_broydn", "scftb_broydn_device", "scftb_broydn_device_ex", "scftb_adm_chen_batch", "scftb_adm_batch", "scftb_adm_mixer_create",
           "scftb_set_diblock", "scftb_residual_ab", "scftb_residual_ab_batch", "scftb_get_phi_ab", "scftb_callback_ab_c0",
           "scftb_mixer_create", "scftb_mixer_destroy", "scftb_mixer_reset", "scftb_mixer_iterate_device",
           "scftb_mixer_status", "scftb_mixer_get_x", "scftb_mixer_get_y", "scftb_get_slots", "scftb_get_kernel_name", "scftb_mixer_set_freeze", "scftb_set_timing", "scftb_get_march_ms", "scftb_spline", "scftb_refine_mesh", "scftb_refine_mesh_adaptive", "scftb_write_solution",
           "scftb_read_solution", "scftb_write_detailed_solution", "scftb_read_res", "scftb2d_nccl_unique_id", "scftb2d_create", "scftb2d_destroy",
           "scftb2d_rows", "scftb_pmixer_create", "scftb_pmixer_destroy", "scftb_pmixer_reset", "scftb_pmixer_iterate_device",
           "scftb_pmixer_status", "scftb_pmixer_get_x", "scftb_padm_batch", "scftb_refine_uniform_batch_device",
           "scftb_free_energy_weights", "scftb_sweep_create", "scftb_sweep_destroy", "scftb_sweep_target_N", "scftb_sweep_solve",
           "scftb2d_p2p_handle", "scftb2d_p2p_attach", "scftb2d_p2p_detach", "scftb2d_residual", "scftb2d_get_phi", "scftb2d_get_stats", "scftb2d_export_csr"]


def lib():
    """Load the C-ABI library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScftError(f"{LIB_PATH} is missing: build it with `make -C scft_b200/csrc` "
                            "(or python -c 'import __graft_entry__ as g; g.build()')")
        L = C.CDLL(LIB_PATH)
        L.scftb_last_error.restype = C.c_char_p
        L.scftb_launch_count.restype = C.c_long
        L.scftb_launch_count.argtypes = [C.c_int]
        L.scftb_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
        L.scftb_destroy.argtypes = [C.c_void_p]
        L.scftb_set_problem.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp]
        L.scftb_residual.argtypes = [C.c_void_p, _dp, _dp]
        L.scftb_residual_batch.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.scftb_residual_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        for nm in ("scftb_get_phi", "scftb_get_f0_given", "scftb_get_eta_full", "scftb_get_q_history", "scftb_get_Q"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, _dp]
        L.scftb_free_energy.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp]
        L.scftb_bind_global.argtypes = [C.c_void_p]
        L.scftb_set_diblock.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.scftb_residual_ab.argtypes = [C.c_void_p, _dp, _dp]
        L.scftb_residual_ab_batch.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.scftb_get_phi_ab.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        for nm in ("scftb_callback_nr1", "scftb_callback_c0", "scftb_callback_fixedpoint_c0", "scftb_callback_ab_c0"):
            getattr(L, nm).restype = None
        L.scftb_adm_chen.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.scftb_adm.argtypes = [C.c_void_p, _dp, C.c_int, _ip, C.c_int]
        L.scftb_broydn.argtypes = [C.c_void_p, _dp, C.c_int, _ip, _dp, _ip]
        L.scftb_broydn_device.argtypes = [C.c_void_p, _dp, _ip, _dp, _ip]
        L.scftb_broydn_device_ex.argtypes = [C.c_void_p, _dp, _ip, _dp, _ip, C.c_int]
        L.scftb_adm_chen_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int, C.c_double, C.c_int,
                                           C.c_int, _ip, _dp]
        L.scftb_mixer_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                         C.POINTER(C.c_void_p)]
        L.scftb_adm_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _ip, _dp]
        L.scftb_adm_mixer_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.scftb_mixer_destroy.argtypes = [C.c_void_p]
        L.scftb_mixer_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.scftb_mixer_iterate_device.argtypes = [C.c_void_p, C.c_void_p]
        L.scftb_mixer_set_freeze.argtypes = [C.c_void_p, C.c_int]
        L.scftb_mixer_status.argtypes = [C.c_void_p, C.c_void_p, _ip, _ip, _dp]
        L.scftb_mixer_get_x.argtypes = [C.c_void_p, C.c_void_p, _dp]
        L.scftb_mixer_get_y.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp]
        L.scftb_get_slots.argtypes = [C.c_void_p, _ip]
        L.scftb_get_kernel_name.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.scftb_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.scftb_get_march_ms.argtypes = [C.c_void_p, _dp, _ip]
        L.scftb_spline.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.scftb_refine_mesh.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.scftb_refine_mesh_adaptive.argtypes = [C.c_int, _dp, _dp, C.c_double, _ip, _dp, _dp]
        L.scftb_write_solution.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.scftb_read_solution.argtypes = [C.c_char_p, _ip, _dp, _dp, C.c_int]
        L.scftb_write_detailed_solution.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int]
        L.scftb_read_res.argtypes = [C.c_char_p, C.c_int, _dp, _dp, _dp]
        L.scftb_pmixer_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_void_p)]
        L.scftb_pmixer_destroy.argtypes = [C.c_void_p]
        L.scftb_pmixer_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.scftb_pmixer_iterate_device.argtypes = [C.c_void_p, C.c_void_p]
        L.scftb_pmixer_status.argtypes = [C.c_void_p, C.c_void_p, _ip, _ip, _dp]
        L.scftb_pmixer_get_x.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.scftb_padm_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int, C.c_int, _ip, _dp]
        L.scftb_refine_uniform_batch_device.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.scftb_free_energy_weights.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp, _dp]
        L.scftb_sweep_create.argtypes = [C.POINTER(_SweepConfig), C.c_int, C.POINTER(C.c_void_p)]
        L.scftb_sweep_destroy.argtypes = [C.c_void_p]
        L.scftb_sweep_target_N.argtypes = [C.c_void_p]
        L.scftb_sweep_solve.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.scftb2d_nccl_unique_id.argtypes = [C.c_char_p]
        L.scftb2d_create.argtypes = [C.POINTER(_Config2D), C.c_char_p, C.POINTER(C.c_void_p)]
        L.scftb2d_destroy.argtypes = [C.c_void_p]
        L.scftb2d_rows.argtypes = [C.c_void_p, _ip, _ip]
        L.scftb2d_p2p_handle.argtypes = [C.c_void_p, C.c_char_p]
        L.scftb2d_p2p_attach.argtypes = [C.c_void_p, C.c_char_p]
        L.scftb2d_p2p_detach.argtypes = [C.c_void_p]
        L.scftb2d_residual.argtypes = [C.c_void_p, _dp, _dp]
        L.scftb2d_get_phi.argtypes = [C.c_void_p, _dp]
        L.scftb2d_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), _dp]
        L.scftb2d_export_csr.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp]
        _lib = L
    return _lib


def launch_count(reset=False):
    return lib().scftb_launch_count(1 if reset else 0)


def _p(a):
    return a.ctypes.data_as(_dp)


def _chk(rc):
    if rc != 0:
        raise ScftError(f"scft_b200 error {rc}: {lib().scftb_last_error().decode()}")


class Engine:
    def __init__(self, N, nsteps=2048, scheme=IE_CONSISTENT, tau=TAU_REF, L=L_REF, quadrature=QUAD_ROMBERG,
                 sign=1.0, max_batch=1, device=0, store_history=False, x=None):
        self.N, self.ni, self.nsteps, self.max_batch = N, N - 2, nsteps, max_batch
        cfg = _Config(scheme, N, nsteps, quadrature, sign, max_batch, device, int(store_history))
        h = C.c_void_p()
        _chk(lib().scftb_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.set_problem(-1, tau, L, x)

    def close(self):
        if getattr(self, "_h", None):
            lib().scftb_destroy(self._h)
            self._h = None

    __del__ = close

    def set_problem(self, p, tau, L, x=None):
        xa = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        _chk(lib().scftb_set_problem(self._h, p, tau, L, _p(xa) if xa is not None else None))

    def residual(self, eta_mid):
        """host eta_mid [ni] or [nprob, ni] -> residual of the same shape (H2D + kernel + D2H)."""
        eta = np.ascontiguousarray(eta_mid, dtype=np.float64)
        out = np.empty_like(eta)
        nprob = 1 if eta.ndim == 1 else eta.shape[0]
        assert eta.shape[-1] == self.ni
        _chk(lib().scftb_residual_batch(self._h, nprob, _p(eta), _p(out)))
        return out

    run = residual  # HeatEquation<dim>::run (drivescft.cc:81)

    def residual_host_ptr(self, nprob, h_eta_ptr, h_out_ptr):
        """scftb_residual_batch on caller-owned (e.g. pinned) host buffers given as addresses"""
        _chk(lib().scftb_residual_batch(self._h, nprob, C.cast(h_eta_ptr, _dp), C.cast(h_out_ptr, _dp)))

    def set_timing(self, on=True):
        _chk(lib().scftb_set_timing(self._h, int(on)))

    def march_ms(self):
        """(total device ms, launches) of the march kernel since the last call"""
        tot, cnt = C.c_double(0), C.c_int(0)
        _chk(lib().scftb_get_march_ms(self._h, C.byref(tot), C.byref(cnt)))
        return tot.value, cnt.value

    def slots(self):
        """resident CTA slots of the march kernel (problems per wave)"""
        v = C.c_int(0)
        _chk(lib().scftb_get_slots(self._h, C.byref(v)))
        return v.value

    def kernel_name(self):
        buf = C.create_string_buffer(128)
        _chk(lib().scftb_get_kernel_name(self._h, buf, 128))
        return buf.value.decode()

    def residual_device(self, nprob, d_eta_ptr, d_out_ptr, stream_ptr=0):
        _chk(lib().scftb_residual_batch_device(self._h, nprob, C.c_void_p(d_eta_ptr), C.c_void_p(d_out_ptr),
                                               C.c_void_p(stream_ptr)))

    def _get(self, fn, p, n):
        a = np.empty(n)
        _chk(fn(self._h, p, _p(a)))
        return a

    def phi(self, p=0):
        return self._get(lib().scftb_get_phi, p, self.N)

    def Q(self, p=0):
        return float(self._get(lib().scftb_get_Q, p, 1)[0])

    def f0_given(self, p=0):
        return self._get(lib().scftb_get_f0_given, p, self.N)

    def eta_full(self, p=0):
        return self._get(lib().scftb_get_eta_full, p, self.N)

    def q_history(self, p=0):
        return self._get(lib().scftb_get_q_history, p, self.N * (self.nsteps + 1)).reshape(self.N, self.nsteps + 1)

    def free_energy(self, p=0, f0bar=0.892581217773656):
        F = C.c_double(0)
        _chk(lib().scftb_free_energy(self._h, p, f0bar, C.byref(F)))
        return F.value

    # ---- two-species (AB diblock) extension: q and q+ as separate sweeps
    def set_diblock(self, fA, chiN, p=-1):
        _chk(lib().scftb_set_diblock(self._h, p, fA, chiN))
        self.two_species = True

    def unknowns(self):
        """unknowns per problem: N-2, or 2(N-2) = (eta_A, eta_B) after set_diblock"""
        return 2 * self.ni if getattr(self, "two_species", False) else self.ni

    def residual_ab(self, w):
        """w [2*ni] or [nprob, 2*ni] = (eta_A, eta_B) on the interior nodes -> residual of the same shape:
        (sign*(phi0 - phiA - phiB), eta_A - eta_B - chiN*(phiB - phiA))"""
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.empty_like(w)
        nprob = 1 if w.ndim == 1 else w.shape[0]
        assert w.shape[-1] == 2 * self.ni
        _chk(lib().scftb_residual_ab_batch(self._h, nprob, _p(w), _p(out)))
        return out

    def phi_ab(self, p=0):
        a, b = np.empty(self.N), np.empty(self.N)
        _chk(lib().scftb_get_phi_ab(self._h, p, _p(a), _p(b)))
        return a, b

    def broydn_device(self, x0, tolf, jc=0, keep_trial=False):
        """scftb_broydn_device[_ex]: (rc, check, x, err, jc); keep_trial = SCFTB_BROYDN_KEEP_TRIAL"""
        x = np.ascontiguousarray(x0, dtype=np.float64).copy()
        chk, err, jcv = C.c_int(1), C.c_double(tolf), C.c_int(jc)
        rc = lib().scftb_broydn_device_ex(self._h, _p(x), C.byref(chk), C.byref(err), C.byref(jcv), int(keep_trial))
        if rc not in (0, 4):
            _chk(rc)
        return rc, chk.value, x, err.value, jcv.value

    def bind_global(self):
        _chk(lib().scftb_bind_global(self._h))

    def adm_chen_batch(self, x, tol, max_iteration, lmd, nn, final=False):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        nprob = 1 if x.ndim == 1 else x.shape[0]
        iters = np.zeros(nprob, dtype=np.int32)
        err = np.zeros(nprob)
        rc = lib().scftb_adm_chen_batch(self._h, nprob, _p(x), tol, max_iteration, lmd, nn, int(final),
                                        iters.ctypes.data_as(_ip), _p(err))
        if rc not in (0, 4):
            _chk(rc)
        return rc, x, iters, err


def _adm_batch(self, x, maxits):
    """scftb_adm_batch: (rc, x, iteration index per problem, err)"""
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    nprob = 1 if x.ndim == 1 else x.shape[0]
    iters = np.zeros(nprob, dtype=np.int32)
    err = np.zeros(nprob)
    rc = lib().scftb_adm_batch(self._h, nprob, _p(x), maxits, iters.ctypes.data_as(_ip), _p(err))
    if rc not in (0, 3, 4):
        _chk(rc)
    return rc, x, iters, err


Engine.adm_batch = _adm_batch


class AndersonBatch:
    """Device-resident Anderson mixing of a batch (scftb_mixer_*): adm_chen per problem (adm=True: adm.c semantics)."""

    def __init__(self, eng, nprob, tol=1e-7, lmd=0.9, nn=3, final=False, adm=False):
        self.eng, self.nprob = eng, nprob
        h = C.c_void_p()
        if adm:
            _chk(lib().scftb_adm_mixer_create(eng._h, nprob, C.byref(h)))
        else:
            _chk(lib().scftb_mixer_create(eng._h, nprob, tol, lmd, nn, int(final), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().scftb_mixer_destroy(self._h)
            self._h = None

    __del__ = close

    def set_freeze(self, freeze):
        _chk(lib().scftb_mixer_set_freeze(self._h, int(freeze)))

    def reset(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        _chk(lib().scftb_mixer_reset(self._h, C.c_void_p(x.ctypes.data), 0, None))

    def reset_device(self, d_x_ptr, stream_ptr=0):
        _chk(lib().scftb_mixer_reset(self._h, C.c_void_p(d_x_ptr), 1, C.c_void_p(stream_ptr)))

    def iterate_device(self, stream_ptr=0):
        _chk(lib().scftb_mixer_iterate_device(self._h, C.c_void_p(stream_ptr)))

    def status(self, stream_ptr=0):
        done = np.zeros(self.nprob, dtype=np.int32)
        iters = np.zeros(self.nprob, dtype=np.int32)
        err = np.zeros(self.nprob)
        _chk(lib().scftb_mixer_status(self._h, C.c_void_p(stream_ptr), done.ctypes.data_as(_ip),
                                      iters.ctypes.data_as(_ip), _p(err)))
        return done, iters, err

    def x(self, stream_ptr=0):
        out = np.zeros((self.nprob, self.eng.unknowns()))
        _chk(lib().scftb_mixer_get_x(self._h, C.c_void_p(stream_ptr), _p(out)))
        return out

    def y(self, stream_ptr, k):
        """residuals F(X_k) of iteration k"""
        out = np.zeros((self.nprob, self.eng.unknowns()))
        _chk(lib().scftb_mixer_get_y(self._h, C.c_void_p(stream_ptr), k, _p(out)))
        return out


class PrecondAndersonBatch:
    """Device-resident preconditioned Anderson mixing of a batch (scftb_pmixer_*)."""

    def __init__(self, eng, nprob, tol=1e-9, nn=10, cap=2.0):
        self.eng, self.nprob = eng, nprob
        h = C.c_void_p()
        _chk(lib().scftb_pmixer_create(eng._h, nprob, tol, nn, cap, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().scftb_pmixer_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        _chk(lib().scftb_pmixer_reset(self._h, C.c_void_p(x.ctypes.data), 0, None))

    def reset_device(self, d_x_ptr, stream_ptr=0):
        _chk(lib().scftb_pmixer_reset(self._h, C.c_void_p(d_x_ptr), 1, C.c_void_p(stream_ptr)))

    def iterate_device(self, stream_ptr=0):
        _chk(lib().scftb_pmixer_iterate_device(self._h, C.c_void_p(stream_ptr)))

    def status(self, stream_ptr=0):
        done = np.zeros(self.nprob, dtype=np.int32)
        iters = np.zeros(self.nprob, dtype=np.int32)
        err = np.zeros(self.nprob)
        _chk(lib().scftb_pmixer_status(self._h, C.c_void_p(stream_ptr), done.ctypes.data_as(_ip),
                                       iters.ctypes.data_as(_ip), _p(err)))
        return done, iters, err

    def x(self, stream_ptr=0):
        out = np.zeros((self.nprob, self.eng.ni))
        _chk(lib().scftb_pmixer_get_x(self._h, C.c_void_p(stream_ptr), C.c_void_p(out.ctypes.data), 0))
        return out


def padm_batch(eng, x, tol=1e-9, max_iteration=200, nn=10):
    """scftb_padm_batch: (rc, x, evaluations-1 per problem, err)"""
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    nprob = 1 if x.ndim == 1 else x.shape[0]
    iters = np.zeros(nprob, dtype=np.int32)
    err = np.zeros(nprob)
    rc = lib().scftb_padm_batch(eng._h, nprob, _p(x), tol, max_iteration, nn, iters.ctypes.data_as(_ip), _p(err))
    if rc not in (0, 3, 4):
        _chk(rc)
    return rc, x, iters, err


def free_energy_weights(N, tau, L, x=None):
    """scftb_free_energy_weights: (c[N], f0bar)"""
    c, f0bar = np.zeros(N), C.c_double(0)
    xa = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    _chk(lib().scftb_free_energy_weights(N, _p(xa) if xa is not None else None, tau, L, _p(c), C.byref(f0bar)))
    return c, f0bar.value


class SweepSolver:
    """scftb_sweep_*: continuation N0 -> ... -> N_target with preconditioned mixing for a batch of problems."""

    def __init__(self, max_prob, N0=33, levels=6, nsteps=2048, scheme=IE_ROWSCALE, quadrature=QUAD_ROMBERG, tol=1e-9, nn=10,
                 maxit=200, cap=2.0, device=0):
        cfg = _SweepConfig(scheme, N0, levels, nsteps, quadrature, tol, nn, maxit, cap, device)
        h = C.c_void_p()
        _chk(lib().scftb_sweep_create(C.byref(cfg), max_prob, C.byref(h)))
        self._h, self.N0, self.levels, self.max_prob = h, N0, levels, max_prob
        self.N_target = lib().scftb_sweep_target_N(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().scftb_sweep_destroy(self._h)
            self._h = None

    __del__ = close

    def solve(self, taus, Ls, eta0, want_fields=True):
        """-> dict(rows [nprob, 7] (status, err, evals, Q, F, evals_last_level, N_last), eta [nprob, N_target-2], level_seconds)"""
        taus, Ls, eta0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (taus, Ls, eta0))
        nprob = len(taus)
        assert eta0.shape == (nprob, self.N0 - 2)
        rows = np.zeros((nprob, SWEEP_COLS))
        eta = np.zeros((nprob, self.N_target - 2)) if want_fields else None
        secs = np.zeros(self.levels + 1)
        _chk(lib().scftb_sweep_solve(self._h, nprob, _p(taus), _p(Ls), _p(eta0), _p(eta) if want_fields else None, _p(rows), _p(secs)))
        return dict(rows=rows, eta=eta, level_seconds=secs)


def spline(x, y, xp, mode=0, bc=0.0):
    """scftb_spline: mode 0 natural, 1 not-a-knot, 2 given second derivative"""
    x, y, xp = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, xp))
    yp = np.zeros_like(xp)
    _chk(lib().scftb_spline(_p(x), _p(y), _p(xp), _p(yp), len(x), len(xp), mode, bc))
    return yp


def refine_mesh(x, eta_mid):
    """scftb_refine_mesh: (x_new[2N-1], eta_mid_new[2N-3])"""
    x, eta_mid = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(eta_mid, dtype=np.float64)
    N = len(x)
    xn, en = np.zeros(2 * N - 1), np.zeros(2 * N - 3)
    _chk(lib().scftb_refine_mesh(N, _p(x), _p(eta_mid), _p(xn), _p(en)))
    return xn, en


def refine_mesh_adaptive(x, eta_mid, factor=10.0):
    """scftb_refine_mesh_adaptive (Matlab_files/refine_mesh.m): (x_new, eta_mid_new) on a locally bisected mesh"""
    x, eta_mid = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(eta_mid, dtype=np.float64)
    N = len(x)
    xn, en = np.zeros(2 * N - 1), np.zeros(2 * N - 3)
    nn = C.c_int(0)
    _chk(lib().scftb_refine_mesh_adaptive(N, _p(x), _p(eta_mid), factor, C.byref(nn), _p(xn), _p(en)))
    return xn[: nn.value].copy(), en[: nn.value - 2].copy()


def write_solution(path, err, F, x, eta_full):
    x, eta_full = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(eta_full, dtype=np.float64)
    _chk(lib().scftb_write_solution(path.encode(), len(x), err, F, _p(x), _p(eta_full)))


def write_detailed_solution(path, err, F, x, eta_full, nplot=0):
    x, eta_full = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(eta_full, dtype=np.float64)
    _chk(lib().scftb_write_detailed_solution(path.encode(), len(x), err, F, _p(x), _p(eta_full), nplot))


def read_solution(path):
    n = C.c_int(0)
    _chk(lib().scftb_read_solution(path.encode(), C.byref(n), None, None, 0))
    x, eta = np.zeros(n.value), np.zeros(n.value)
    _chk(lib().scftb_read_solution(path.encode(), C.byref(n), _p(x), _p(eta), n.value))
    return x, eta


def read_res(path, rows):
    xl, phi, eta = np.zeros(rows), np.zeros(rows), np.zeros(rows)
    _chk(lib().scftb_read_res(path.encode(), rows, _p(xl), _p(phi), _p(eta)))
    return xl, phi, eta


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _chk(lib().scftb2d_nccl_unique_id(buf))
    return buf.raw


class Engine2D:
    """2-D Q1 mesh, CSR matrices, Jacobi-PCG per contour step (scftb2d_*).  world > 1: slab partition
    in x; pass the 128-byte NCCL id obtained on rank 0 (nccl_unique_id()) to every rank."""

    def __init__(self, nx, ny, L=L_REF, Ly=None, tau=TAU_REF, nsteps=64, quadrature=QUAD_ROMBERG, sign=1.0,
                 rtol=1e-12, maxit=0, device=0, rank=0, world=1, nccl_id=None, store_history=False):
        Ly = L / nx * ny if Ly is None else Ly
        self.nx, self.ny, self.ndof = nx, ny, (nx + 1) * (ny + 1)
        cfg = _Config2D(nx, ny, L, Ly, tau, nsteps, quadrature, sign, rtol, maxit, device, rank, world, int(store_history))
        h = C.c_void_p()
        _chk(lib().scftb2d_create(C.byref(cfg), nccl_id, C.byref(h)))
        self._h = h
        r0, nr = C.c_int(0), C.c_int(0)
        _chk(lib().scftb2d_rows(h, C.byref(r0), C.byref(nr)))
        self.row0, self.nrows = r0.value, nr.value

    def close(self):
        if getattr(self, "_h", None):
            lib().scftb2d_destroy(self._h)
            self._h = None

    __del__ = close

    def p2p_handle(self):
        buf = C.create_string_buffer(64)
        _chk(lib().scftb2d_p2p_handle(self._h, buf))
        return buf.raw

    def p2p_attach(self, handles):
        """handles: bytes of length 64*world (rank order)"""
        _chk(lib().scftb2d_p2p_attach(self._h, handles))

    def p2p_detach(self):
        _chk(lib().scftb2d_p2p_detach(self._h))

    def residual(self, eta):
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        assert eta.size == self.ndof
        out = np.zeros(self.nrows)
        _chk(lib().scftb2d_residual(self._h, _p(eta), _p(out)))
        return out

    def phi(self):
        a = np.zeros(self.nrows)
        _chk(lib().scftb2d_get_phi(self._h, _p(a)))
        return a

    def stats(self):
        it, ms = C.c_longlong(0), C.c_double(0)
        _chk(lib().scftb2d_get_stats(self._h, C.byref(it), C.byref(ms)))
        return it.value, ms.value

    def csr(self):
        rowptr = np.zeros(self.nrows + 1, dtype=np.int32)
        colind = np.zeros(9 * self.nrows, dtype=np.int32)
        vt, va = np.zeros(9 * self.nrows), np.zeros(9 * self.nrows)
        _chk(lib().scftb2d_export_csr(self._h, rowptr.ctypes.data_as(_ip), colind.ctypes.data_as(_ip), _p(vt), _p(va)))
        nnz = rowptr[-1]
        return rowptr, colind[:nnz], vt[:nnz], va[:nnz]
