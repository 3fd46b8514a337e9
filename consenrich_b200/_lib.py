"""ctypes binding of libconsenrich_b200.so (include/consenrich_b200.h).

Loading fails loudly when the library has not been built: there is no CPU implementation to
fall back to.  Creating a context fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import bisect
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libconsenrich_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
ABI_VERSION = 3

FAM_FOLD, FAM_FORWARD, FAM_BACKWARD, FAM_RESIDUALS, FAM_PRECISION = range(5)
FAMILY_NAMES = ("fold", "forward_scan", "backward_scan", "residuals", "precision_updates", "background", "munc",
                "forward_compose", "segment_scan", "backward_publish", "writer")


class NativeLibraryMissing(RuntimeError):
    pass


class CudaError(RuntimeError):
    pass


class Model(C.Structure):
    """cb200_model"""
    _fields_ = [
        ("state_dim", C.c_int32), ("use_lambda", C.c_int32), ("use_kappa", C.c_int32), ("use_qscale", C.c_int32),
        ("return_nll", C.c_int32), ("store_nll_in_d", C.c_int32), ("use_apn", C.c_int32), ("reserved1", C.c_int32),
        ("F", C.c_double * 4), ("Q0", C.c_double * 4),
        ("state_init", C.c_double), ("cov_init", C.c_double), ("pad", C.c_double),
        ("lam_min", C.c_double), ("lam_max", C.c_double), ("kap_min", C.c_double), ("kap_max", C.c_double),
        ("apn_min_q", C.c_double), ("apn_max_q", C.c_double), ("apn_thresh", C.c_double), ("apn_scale", C.c_double),
        ("apn_pc", C.c_double),
    ]


class EcmOpts(C.Structure):
    """cb200_ecm_opts"""
    _fields_ = [
        ("max_iters", C.c_int32), ("inner_iters", C.c_int32), ("update_lambda", C.c_int32),
        ("update_kappa", C.c_int32), ("want_outputs", C.c_int32), ("init_ones", C.c_int32),
        ("rtol", C.c_double), ("nu", C.c_double),
    ]


class EcmResult(C.Structure):
    """cb200_ecm_result"""
    _fields_ = [
        ("iters_done", C.c_int32), ("converged", C.c_int32), ("skipped", C.c_int32), ("stable_iters", C.c_int32),
        ("nll_increase_count", C.c_int32), ("has_initial", C.c_int32),
        ("initial_nll", C.c_double), ("final_nll", C.c_double), ("final_abs_rel_change", C.c_double),
        ("final_rel_improvement", C.c_double),
    ]


class MuncFinalizeResult(C.Structure):
    """cb200_munc_finalize_result"""
    _fields_ = [(k, C.c_int64) for k in ("support_count", "count_floor_finite", "count_floor_added",
                                         "count_floor_missing", "invalid_local", "invalid_prior",
                                         "invalid_count_floor")]


class DiagGainArgs(C.Structure):
    """cb200_diag_gain_args"""
    _fields_ = ([(k, C.c_void_p) for k in ("covar", "p_noise", "q_scale", "proc_prec", "sum_inv_r", "sum_gain0",
                                           "sum_gain1")]
                + [("n", C.c_int64), ("dim", C.c_int32), ("cov_dim", C.c_int32), ("base_q", C.c_double * 4),
                   ("f", C.c_double * 4), ("cov_init", C.c_double)])


class MuncSeedArgs(C.Structure):
    """cb200_munc_seed_args"""
    _fields_ = ([(k, C.c_void_p) for k in ("data", "munc", "state_mean", "state_var", "background", "g_var",
                                           "count_floor", "omega_in", "rho_in", "active", "moment", "rho_out",
                                           "omega_raw", "omega_out", "local", "variance")]
                + [(k, C.c_int64) for k in ("m", "n", "ld", "active_ld")]
                + [(k, C.c_int32) for k in ("active_mode", "use_weights", "student_t", "update_weights")]
                + [(k, C.c_double) for k in ("pad", "student_t_df", "d_omega", "omega_min", "omega_max",
                                             "variance_floor", "variance_cap")])


_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t
_pm, _po, _pr = C.POINTER(Model), C.POINTER(EcmOpts), C.POINTER(EcmResult)

# name -> (restype, argtypes); one entry per declaration in include/consenrich_b200.h
SIGNATURES = {
    "cb200_abi_version": (C.c_int, []),
    "cb200_current_device": (C.c_int, [C.POINTER(C.c_int)]),
    "cb200_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "cb200_ctx_destroy": (None, [_vp]),
    "cb200_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "cb200_ctx_sync": (C.c_int, [_vp]),
    "cb200_last_error": (C.c_char_p, []),
    "cb200_ctx_launch_count": (_i64, [_vp]),
    "cb200_ctx_enable_timing": (C.c_int, [_vp, C.c_int]),
    "cb200_ctx_kernel_ms": (C.c_int, [_vp, C.c_int, C.POINTER(_dbl), C.POINTER(_i64)]),
    "cb200_ctx_reset_timing": (C.c_int, [_vp]),
    "cb200_ctx_set_timing_stride": (C.c_int, [_vp, C.c_int]),
    "cb200_ctx_kernel_launches": (C.c_int, [_vp, C.c_int, C.POINTER(_i64)]),
    "cb200_set_scan_substeps": (C.c_int, [C.c_int]),
    "cb200_set_lean_sweeps": (C.c_int, [C.c_int, C.c_int]),
    "cb200_debug_scan_times": (C.c_int, [_vp, _i64, _vp]),
    "cb200_device_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "cb200_device_free": (C.c_int, [_vp, _vp]),
    "cb200_pinned_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "cb200_pinned_free": (C.c_int, [_vp]),
    "cb200_copy_h2d": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "cb200_copy_d2h": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "cb200_fold_tracks": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _dbl, _vp, _i64]),
    "cb200_forward_scan": (C.c_int, [_vp, _pm, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cb200_forward_scan_shard": (C.c_int, [_vp, _pm, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp]),
    "cb200_forward_shard_aggregate": (C.c_int, [_vp, _pm, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "cb200_forward_shard_prefix": (C.c_int, [_vp, _pm, _vp, _i32, _vp]),
    "cb200_backward_scan": (C.c_int, [_vp, _pm, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "cb200_backward_scan_kappa": (C.c_int, [_vp, _pm, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _dbl, _vp]),
    "cb200_backward_shard_aggregate": (C.c_int, [_vp, _pm, _i64, _vp, _vp, _vp, _i32, _vp]),
    "cb200_backward_shard_prefix": (C.c_int, [_vp, _pm, _vp, _i32, _i32, _vp]),
    "cb200_residuals": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp]),
    "cb200_update_lambda": (C.c_int, [_vp, _pm, _vp, _i64, _i64, _i64, _vp, _vp, _dbl, _vp]),
    "cb200_update_kappa": (C.c_int, [_vp, _pm, _i64, _vp, _vp, _vp, _vp, _dbl, _vp]),
    "cb200_ecm_device": (C.c_int, [_vp, _pm, _po, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _pr, _vp]),
    "cb200_host_forward_pass": (C.c_int, [_vp, _pm, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "cb200_host_backward_pass": (C.c_int, [_vp, _pm, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cb200_host_sweep": (C.c_int, [_vp, _pm, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   C.POINTER(_dbl), C.POINTER(_dbl), _vp, _vp, _vp, _i64, _vp]),
    "cb200_host_ecm": (C.c_int, [_vp, _pm, _po, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _pr, _vp]),
    "cb200_background_stats": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "cb200_background_solve": (C.c_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _i32, _vp, C.POINTER(_i64), C.POINTER(_dbl)]),
    "cb200_host_background_stats": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, C.POINTER(_i64)]),
    "cb200_host_background_solve": (C.c_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _i32, _vp, C.POINTER(_i64),
                                              C.POINTER(_dbl)]),
    "cb200_munc_smooth_local_evidence": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _dbl, _vp, _i64,
                                                   _vp]),
    "cb200_host_munc_smooth_local_evidence": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _dbl, _vp,
                                                        C.POINTER(_i32)]),
    "cb200_munc_finalize_eb": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _vp,
                                         C.POINTER(MuncFinalizeResult)]),
    "cb200_host_munc_finalize_eb": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _vp,
                                              C.POINTER(MuncFinalizeResult)]),
    "cb200_munc_seed_pass": (C.c_int, [_vp, C.POINTER(MuncSeedArgs), _vp]),
    "cb200_host_munc_seed_pass": (C.c_int, [_vp, C.POINTER(MuncSeedArgs), C.POINTER(_i32)]),
    "cb200_weighted_mean_residual": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _dbl, _vp]),
    "cb200_host_weighted_mean_residual": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _dbl, _vp]),
    "cb200_diag_obs_sums": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _dbl, _vp, _vp]),
    "cb200_diag_gain": (C.c_int, [_vp, C.POINTER(DiagGainArgs)]),
    "cb200_host_interval_diagnostics": (C.c_int, [_vp, _vp, _i64, _vp, _dbl, C.POINTER(DiagGainArgs), _vp, _vp]),
    "cb200_ema": (C.c_int, [_vp, _vp, _i64, _i32, _dbl, _vp]),
    "cb200_host_ema": (C.c_int, [_vp, _vp, _i64, _i32, _dbl, _vp]),
    "cb200_split_begin": (C.c_int, [_vp, _pm, _dbl, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i32, _i32]),
    "cb200_split_forward_compose": (C.c_int, [_vp, _vp]),
    "cb200_split_forward_replay": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "cb200_split_backward_compose": (C.c_int, [_vp, _i32, _vp]),
    "cb200_split_backward_replay": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cb200_split_end": (C.c_int, [_vp, _vp]),
    "cb200_bedgraph_chunk": (C.c_int, [_vp, C.c_char_p, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _i64, C.POINTER(_vp),
                                       C.POINTER(_i64)]),
    "cb200_host_bedgraph_chunk": (C.c_int, [_vp, C.c_char_p, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _i64,
                                            C.POINTER(_vp), C.POINTER(_i64)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libconsenrich_b200.so; raises NativeLibraryMissing if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} is missing. Build it with `python -m consenrich_b200.build` (needs nvcc); "
                "consenrich_b200 has no CPU implementation to fall back to.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.cb200_abi_version() != ABI_VERSION:
            raise NativeLibraryMissing(f"{LIB_PATH} has ABI {lib.cb200_abi_version()}, expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def last_error() -> str:
    msg = load().cb200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Map a C status to the reference's exception types (ValueError for argument errors)."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CudaError(msg)


class Context:
    """One cb200_ctx: a device, a stream, the device arena and launch accounting."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load()
        h = _vp()
        check(self._lib.cb200_ctx_create(int(device), _vp(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.stream_handle = int(stream) if stream else 0  # 0: a stream owned by the context

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cb200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream: int | None):
        check(self._lib.cb200_ctx_set_stream(self.handle, _vp(stream) if stream else None))
        self.stream_handle = int(stream) if stream else 0

    def sync(self):
        check(self._lib.cb200_ctx_sync(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._lib.cb200_ctx_launch_count(self.handle))

    def enable_timing(self, on: bool = True, stride: int = 1):
        """Bracket every ``stride``-th launch of each kernel family with CUDA events."""
        check(self._lib.cb200_ctx_set_timing_stride(self.handle, int(stride)))
        check(self._lib.cb200_ctx_enable_timing(self.handle, int(on)))

    def reset_timing(self):
        check(self._lib.cb200_ctx_reset_timing(self.handle))

    def kernel_ms(self) -> dict:
        out = {}
        for fam, name in enumerate(FAMILY_NAMES):
            ms, cnt = _dbl(), _i64()
            check(self._lib.cb200_ctx_kernel_ms(self.handle, fam, C.byref(ms), C.byref(cnt)))
            out[name] = (ms.value, cnt.value)
        return out

    def kernel_launches(self) -> dict:
        """Launches of each family made while timing was enabled (all of them, not only the sampled ones)."""
        out = {}
        for fam, name in enumerate(FAMILY_NAMES):
            cnt = _i64()
            check(self._lib.cb200_ctx_kernel_launches(self.handle, fam, C.byref(cnt)))
            out[name] = cnt.value
        return out


# ------------------------------------------------------------------------------------------
# pinned result arrays
# ------------------------------------------------------------------------------------------
# The arrays the drop-in functions hand back are allocated in page-locked host memory, so that the
# device -> host copy lands in them directly at full PCIe rate (a pageable destination is staged
# through a bounce buffer by the driver at a fraction of that).  Blocks are recycled through a
# small pool when the numpy array that wraps them is garbage collected: cudaHostAlloc is far too
# slow to call per result.
_pin_pool: list = []      # free blocks, (nbytes, ptr), sorted by size
_pin_lock = threading.Lock()
_PIN_MIN = 1 << 16        # smaller results stay ordinary numpy arrays
_PIN_POOL_CAP = int(os.environ.get("CB200_PINNED_POOL_BYTES", 24 << 30))   # bytes kept for reuse


class _PinnedBlock:
    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def __del__(self):
        try:
            with _pin_lock:
                held = sum(k for k, _ in _pin_pool)
                if held + self.nbytes <= _PIN_POOL_CAP:
                    bisect.insort(_pin_pool, (self.nbytes, self.ptr))
                    return
            load().cb200_pinned_free(_vp(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked memory (falls back to numpy.empty for small results).  A freed block
    serves any later request it is large enough for (best fit, at most 4x the request): a genome's
    chromosomes come in decreasing sizes, and cudaHostAlloc costs ~0.3 s per GB."""
    import numpy as np
    dt = np.dtype(dtype)
    count = int(np.prod(shape)) if len(shape) else 1
    nbytes = count * dt.itemsize
    if nbytes < _PIN_MIN:
        return np.empty(shape, dt)
    size = 1 << (nbytes - 1).bit_length() if nbytes < (1 << 24) else (nbytes + (1 << 22) - 1) >> 22 << 22
    ptr = None
    with _pin_lock:
        i = bisect.bisect_left(_pin_pool, (size, 0))
        if i < len(_pin_pool) and _pin_pool[i][0] <= 4 * size:
            size, ptr = _pin_pool.pop(i)
    if ptr is None:
        p = _vp()
        check(load().cb200_pinned_alloc(size, C.byref(p)))
        ptr = p.value
    block = _PinnedBlock(ptr, size)
    buf = (C.c_char * nbytes).from_address(ptr)
    buf._cb200_block = block  # the block lives exactly as long as the buffer numpy keeps as its base
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


_default_ctx = threading.local()  # per thread: a context (stream + device arena) serves one call at a time


def current_device() -> int:
    d = C.c_int(0)
    check(load().cb200_current_device(C.byref(d)))
    return int(d.value)


def default_context(device: int | None = None) -> Context:
    """The calling thread's context for a device, used by the drop-in functions; the device defaults to
    the one the thread has selected (cudaGetDevice), i.e. the rank's GPU under torchrun.  Contexts are per
    thread because the reference's functions are re-entrant and its MUNC stage calls them from a thread
    pool (consenrich.py:9055): each thread gets its own stream and device buffers."""
    if device is None:
        device = current_device()
    table = getattr(_default_ctx, "by_device", None)
    if table is None:
        table = _default_ctx.by_device = {}
    ctx = table.get(device)
    if ctx is None:
        ctx = table[device] = Context(device)
    return ctx
