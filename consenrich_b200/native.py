"""Host-side mirror of the six native entry points of ``consenrich.cconsenrich`` that sit on the
state-space hot path, backed by libconsenrich_b200.so (sm_100a).

Same names, keyword arguments, return tuples and ValueError texts as the reference
(``/root/reference/src/consenrich/cconsenrich.pyx``):

* ``cforwardPass``             pyx:6393-6632      * ``cbackwardPass``            pyx:6635-6850
* ``cforwardPassLevel``        pyx:6853-7049      * ``cbackwardPassLevel``       pyx:7052-7150
* ``cfixedBackgroundECM``      pyx:7660-8442      * ``cfixedBackgroundECMLevel`` pyx:7153-7657

``install()`` replaces those attributes on the reference's module (the seam the reference's own
tests patch, tests/test_core.py:1317), so ``consenrich.core.runConsenrich`` runs on the GPU
unchanged.  Arguments the reference accepts but never forwards to its loops (``chunkSize``,
``projectStateDuringFiltering``, ``stateLowerBound``/``stateUpperBound``; pyx:6403-6406) are
accepted and ignored here too.  Adaptive process noise (``ECM_useAPN`` without a
``processQScale``; pyx:510-527) is a per-bin nonlinear feedback that no associative scan can
express: that forward pass runs as a sequential recursion on the device (csrc/apn_kernels.cu).

There is no CPU fallback: without the built library or without a CUDA device every function
raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["cforwardPass", "cbackwardPass", "cforwardPassLevel", "cbackwardPassLevel",
           "cfixedBackgroundECM", "cfixedBackgroundECMLevel", "sweep", "install", "uninstall"]

_HOT_PATH = ("cforwardPass", "cbackwardPass", "cforwardPassLevel", "cbackwardPassLevel",
             "cfixedBackgroundECM", "cfixedBackgroundECMLevel")
# the callers on the other side of the ECM inside an outer pass
_BACKGROUND = ("cbackgroundWeightedStats", "cbackgroundWeightedStatsWithSupport", "csolveZeroCenteredBackground")
# dense kernels of the observation-noise stage that produces matrixMunc
_MUNC = ("cMuncSmoothDenseLocalEvidence", "cFinalizeMuncEBTrack", "cMuncObservationMomentSeedPass", "cEMA")


def _f32(x) -> float:
    """Round a Python scalar to C float and widen back (Cython ``float`` arguments)."""
    return float(np.float32(x))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c32(a, name, ndim):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == ndim and a.flags.c_contiguous):
        raise ValueError(f"{name}: Buffer dtype mismatch / not C-contiguous float32 ndim={ndim}")
    return a


def _buf32(a, name, shape_min):
    """A caller-supplied in/out buffer: C-contiguous float32 of at least ``shape_min`` rows and exactly the
    trailing dimensions (the reference's typed memoryviews raise on anything else; a raw pointer would
    read or write out of bounds)."""
    a = _c32(a, name, len(shape_min))
    if a.shape[0] < shape_min[0] or tuple(a.shape[1:]) != tuple(shape_min[1:]):
        raise ValueError(f"{name}: shape {tuple(a.shape)} does not hold {tuple(shape_min)}")
    return a


def _vec32(a, name, n):
    if a is None:
        return None
    a = _c32(a, name, 1)
    if a.shape[0] != n:
        raise ValueError(f"{name} length must match intervalCount")
    return a


def _coerce_qscale(q, n):
    """cconsenrich.pyx:101-131 (_coerceProcessQScale)."""
    arr = np.ascontiguousarray(q, dtype=np.float32).reshape(-1)
    if arr.shape[0] != n:
        raise ValueError("processQScale length must match intervalCount")
    if n > 0:
        a64 = arr.astype(np.float64)
        if (~np.isfinite(a64) | (a64 <= 0.0)).any():
            raise ValueError("processQScale must contain only positive finite values")
        if abs(float(a64[0]) - 1.0) > 1.0e-6:
            raise ValueError("processQScale[0] must be 1.0")
    return arr


def _check_bounds(lo, hi, is_obs):
    """cconsenrich.pyx:143-151."""
    if lo <= 0.0 or hi <= 0.0 or hi < lo:
        raise ValueError(("observation" if is_obs else "process")
                         + " precision multiplier bounds must satisfy 0 < min <= max")


def _apn_live(useAPN, use_qscale, Q0, dim) -> bool:
    """True when the reference would run the adaptive-process-noise feedback (pyx:6574-6576, 510)."""
    if not useAPN or use_qscale:
        return False
    q_diag = 0.5 * (float(Q0[0, 0]) + float(Q0[1, 1])) if dim == 2 else float(Q0[0, 0])
    return q_diag > 1.0e-12


_APN_DEFAULTS = (1.0e-4, 1000.0, 5.0, 10.0, 2.0)  # APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC


def _model(dim, matrixF, Q0, stateInit, stateCovarInit, pad, lamMin, lamMax, kapMin, kapMax, use_lambda, use_kappa,
           use_qscale, returnNLL, storeNLLInD, useAPN=False, apn=_APN_DEFAULTS):
    mo = _lib.Model()
    # adaptive process noise (pyx:510-527): the scalars arrive as C float in the reference (pyx:6422-6426)
    mo.use_apn = int(bool(useAPN))
    mo.apn_min_q, mo.apn_max_q, mo.apn_thresh, mo.apn_scale, mo.apn_pc = map(_f32, apn)
    mo.state_dim = dim
    mo.use_lambda, mo.use_kappa, mo.use_qscale = int(use_lambda), int(use_kappa), int(use_qscale)
    mo.return_nll, mo.store_nll_in_d = int(bool(returnNLL)), int(bool(storeNLLInD))
    if dim == 2:
        Fm = np.asarray(matrixF)
        mo.F[:] = [float(Fm[0, 0]), float(Fm[0, 1]), float(Fm[1, 0]), float(Fm[1, 1])]
        mo.Q0[:] = [float(Q0[0, 0]), float(Q0[0, 1]), float(Q0[1, 0]), float(Q0[1, 1])]
    else:
        mo.F[:] = [1.0, 0.0, 0.0, 1.0]
        mo.Q0[:] = [float(Q0[0, 0]), 0.0, 0.0, 0.0]
    mo.state_init, mo.cov_init, mo.pad = _f32(stateInit), _f32(stateCovarInit), _f32(pad)
    mo.lam_min, mo.lam_max, mo.kap_min, mo.kap_max = lamMin, lamMax, kapMin, kapMax
    return mo


def _ctx(device=None):
    return _lib.default_context(None if device is None else int(device))


def _forward(dim, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount, stateInit,
             stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD, returnNLL, storeNLLInD,
             lambdaExp, processPrecExp, useObs, useProc, useAPN, lamMin, lamMax, kapMin, kapMax, processQScale,
             apn=_APN_DEFAULTS):
    data = _c32(matrixData, "matrixData", 2)
    munc = _c32(matrixPluginMuncInit, "matrixPluginMuncInit", 2)
    m, n = data.shape
    do_store = stateForward is not None
    use_lambda = bool(useObs and (lambdaExp is not None))
    use_qscale = processQScale is not None
    use_kappa = bool(useProc and (processPrecExp is not None) and ((not useAPN) or use_qscale))
    qs = _coerce_qscale(processQScale, n) if use_qscale else None
    if n <= 0 or m <= 0:  # pyx:6494-6501
        d = np.empty(n, dtype=np.float32) if vectorD is None else vectorD
        return (np.float32(0.0), 0, d, 0.0) if returnNLL else (np.float32(0.0), 0, d)
    if blockCount <= 0:
        raise ValueError("blockCount must be positive")
    if munc.shape[0] != m or munc.shape[1] != n:
        raise ValueError("matrixPluginMuncInit shape must match matrixData shape")
    Q0 = np.asarray(matrixQ0)
    if dim == 2:
        Fm = np.asarray(matrixF)
        if Fm.shape[0] < 2 or Fm.shape[1] < 2:
            raise ValueError("matrixF must have at least shape (2, 2)")
        if Q0.shape[0] < 2 or Q0.shape[1] < 2:
            raise ValueError("matrixQ0 must have at least shape (2, 2)")
    else:
        if Q0.shape[0] < 1 or Q0.shape[1] < 1:
            raise ValueError("matrixQ0 must have at least shape (1, 1)")
        if float(Q0[0, 0]) <= 0.0:
            raise ValueError("matrixQ0[0, 0] must be positive")
    lamMin, lamMax, kapMin, kapMax = map(_f32, (lamMin, lamMax, kapMin, kapMax))
    _check_bounds(lamMin, lamMax, True)
    _check_bounds(kapMin, kapMax, False)
    bm = np.ascontiguousarray(intervalToBlockMap, dtype=np.int32)
    if bm.shape[0] < n:
        raise ValueError("intervalToBlockMap length must match intervalCount")
    lam = _vec32(lambdaExp, "lambdaExp", n) if use_lambda else None
    kap = _vec32(processPrecExp, "processPrecExp", n) if use_kappa else None
    if vectorD is None:
        vectorD = np.empty(n, dtype=np.float32)
    elif vectorD.shape[0] < n:
        raise ValueError("vectorD length must match intervalCount")
    _buf32(vectorD, "vectorD", (n,))
    if do_store:
        _buf32(stateForward, "stateForward", (n, dim))
        _buf32(stateCovarForward, "stateCovarForward", (n, dim, dim))
        _buf32(pNoiseForward, "pNoiseForward", (max(n - 1, 1), dim, dim))
    mo = _model(dim, matrixF, Q0, stateInit, stateCovarInit, pad, lamMin, lamMax, kapMin, kapMax, use_lambda,
                use_kappa, use_qscale, returnNLL, storeNLLInD, useAPN, apn)
    sum_d, sum_nll = C.c_double(0.0), C.c_double(0.0)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_forward_pass(
        ctx.handle, C.byref(mo), _ptr(data), _ptr(munc), m, n, _ptr(bm), int(blockCount), _ptr(lam), _ptr(kap),
        _ptr(qs), _ptr(stateForward) if do_store else None, _ptr(stateCovarForward) if do_store else None,
        _ptr(pNoiseForward) if do_store else None, _ptr(vectorD), C.byref(sum_d), C.byref(sum_nll)))
    phi = float(np.float32(sum_d.value / float(n)))
    if returnNLL:
        return (phi, 0, vectorD, sum_nll.value)
    return (phi, 0, vectorD)


def cforwardPass(matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                 stateInit, stateCovarInit, pad=1.0e-4, projectStateDuringFiltering=False,
                 stateLowerBound=0.0, stateUpperBound=0.0, chunkSize=1000000, stateForward=None,
                 stateCovarForward=None, pNoiseForward=None, vectorD=None, returnNLL=False,
                 storeNLLInD=False, lambdaExp=None, processPrecExp=None,
                 ECM_useObsPrecisionReweighting=True, ECM_useProcessPrecisionReweighting=True,
                 ECM_useAPN=False, obsPrecisionMultiplierMin=0.25, obsPrecisionMultiplierMax=4.0,
                 procPrecisionMultiplierMin=0.25, procPrecisionMultiplierMax=4.0, APN_minQ=1.0e-4,
                 APN_maxQ=1000.0, APN_dStatThresh=5.0, APN_dStatScale=10.0, APN_dStatPC=2.0,
                 processQScale=None):
    """2-state forward filter; signature and returns of cconsenrich.pyx:6393-6632."""
    return _forward(2, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                    stateInit, stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD,
                    returnNLL, storeNLLInD, lambdaExp, processPrecExp, ECM_useObsPrecisionReweighting,
                    ECM_useProcessPrecisionReweighting, ECM_useAPN, obsPrecisionMultiplierMin,
                    obsPrecisionMultiplierMax, procPrecisionMultiplierMin, procPrecisionMultiplierMax, processQScale,
                    (APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC))


def cforwardPassLevel(matrixData, matrixPluginMuncInit, matrixQ0, intervalToBlockMap, blockCount,
                      stateInit, stateCovarInit, pad=1.0e-4, chunkSize=1000000, stateForward=None,
                      stateCovarForward=None, pNoiseForward=None, vectorD=None, returnNLL=False,
                      storeNLLInD=False, lambdaExp=None, processPrecExp=None,
                      ECM_useObsPrecisionReweighting=True, ECM_useProcessPrecisionReweighting=True,
                      ECM_useAPN=False, obsPrecisionMultiplierMin=0.25, obsPrecisionMultiplierMax=4.0,
                      procPrecisionMultiplierMin=0.25, procPrecisionMultiplierMax=4.0, APN_minQ=1.0e-4,
                      APN_maxQ=1000.0, APN_dStatThresh=5.0, APN_dStatScale=10.0, APN_dStatPC=2.0,
                      processQScale=None):
    """Level-only forward filter; signature and returns of cconsenrich.pyx:6853-7049."""
    return _forward(1, matrixData, matrixPluginMuncInit, None, matrixQ0, intervalToBlockMap, blockCount,
                    stateInit, stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD,
                    returnNLL, storeNLLInD, lambdaExp, processPrecExp, ECM_useObsPrecisionReweighting,
                    ECM_useProcessPrecisionReweighting, ECM_useAPN, obsPrecisionMultiplierMin,
                    obsPrecisionMultiplierMax, procPrecisionMultiplierMin, procPrecisionMultiplierMax, processQScale,
                    (APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC))


def _backward(dim, matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward, stateSmoothed,
              stateCovarSmoothed, lagCovSmoothed, postFitResiduals):
    data = _c32(matrixData, "matrixData", 2)
    m, n = data.shape
    alloc = _lib.pinned_empty if n > 0 else np.empty
    xs = stateSmoothed if stateSmoothed is not None else alloc((n, dim), np.float32)
    Ps = stateCovarSmoothed if stateCovarSmoothed is not None else alloc((n, dim, dim), np.float32)
    lag = lagCovSmoothed if lagCovSmoothed is not None else alloc((max(n - 1, 1), dim, dim), np.float32)
    res = postFitResiduals if postFitResiduals is not None else alloc((n, m), np.float32)
    if n <= 0:
        return (xs, Ps, lag, res)
    xf = _buf32(stateForward, "stateForward", (n, dim))
    Pf = _buf32(stateCovarForward, "stateCovarForward", (n, dim, dim))
    Qf = _buf32(pNoiseForward, "pNoiseForward", (max(n - 1, 1), dim, dim))
    _buf32(xs, "stateSmoothed", (n, dim))
    _buf32(Ps, "stateCovarSmoothed", (n, dim, dim))
    _buf32(lag, "lagCovSmoothed", (1, dim, dim))
    _buf32(res, "postFitResiduals", (n, m))
    mo = _model(dim, matrixF, np.eye(2), 0.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, False, False, False, False, False)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_backward_pass(
        ctx.handle, C.byref(mo), _ptr(data), m, n, _ptr(xf), _ptr(Pf), _ptr(Qf), _ptr(xs), _ptr(Ps), _ptr(lag),
        int(lag.shape[0]), _ptr(res)))
    return (xs, Ps, lag, res)


def cbackwardPass(matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward, chunkSize=1000000,
                  stateSmoothed=None, stateCovarSmoothed=None, lagCovSmoothed=None, postFitResiduals=None):
    """2-state RTS smoother + lag-one covariance + residuals; cconsenrich.pyx:6635-6850."""
    return _backward(2, matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward, stateSmoothed,
                     stateCovarSmoothed, lagCovSmoothed, postFitResiduals)


def cbackwardPassLevel(matrixData, stateForward, stateCovarForward, pNoiseForward, chunkSize=1000000,
                       stateSmoothed=None, stateCovarSmoothed=None, lagCovSmoothed=None, postFitResiduals=None):
    """Level-only RTS smoother; cconsenrich.pyx:7052-7150."""
    return _backward(1, matrixData, None, stateForward, stateCovarForward, pNoiseForward, stateSmoothed,
                     stateCovarSmoothed, lagCovSmoothed, postFitResiduals)


def sweep(matrixData, matrixPluginMuncInit, matrixF, matrixQ0, stateInit, stateCovarInit, pad=1.0e-4, stateModel=2,
          lambdaExp=None, processPrecExp=None, processQScale=None, returnNLL=True, storeNLLInD=False,
          obsPrecisionMultiplierMin=0.25, obsPrecisionMultiplierMax=4.0, procPrecisionMultiplierMin=0.25,
          procPrecisionMultiplierMax=4.0, wantResiduals=True, out=None):
    """One forward + backward sweep on a single upload (what core._runForwardBackward, core.py:4207,
    obtains from cforwardPass followed by cbackwardPass).  Returns a dict of the reference's arrays.
    ``out`` may hold preallocated (e.g. pinned) arrays under the same keys."""
    dim = int(stateModel)
    data = _c32(matrixData, "matrixData", 2)
    munc = _c32(matrixPluginMuncInit, "matrixPluginMuncInit", 2)
    m, n = data.shape
    if munc.shape != data.shape:
        raise ValueError("matrixPluginMuncInit shape must match matrixData shape")
    Q0 = np.asarray(matrixQ0)
    lamMin, lamMax, kapMin, kapMax = map(_f32, (obsPrecisionMultiplierMin, obsPrecisionMultiplierMax,
                                                 procPrecisionMultiplierMin, procPrecisionMultiplierMax))
    _check_bounds(lamMin, lamMax, True)
    _check_bounds(kapMin, kapMax, False)
    qs = _coerce_qscale(processQScale, n) if processQScale is not None else None
    lam = _vec32(lambdaExp, "lambdaExp", n)
    kap = _vec32(processPrecExp, "processPrecExp", n)
    mo = _model(dim, matrixF, Q0, stateInit, stateCovarInit, pad, lamMin, lamMax, kapMin, kapMax, lam is not None,
                kap is not None, qs is not None, returnNLL, storeNLLInD)
    out = {} if out is None else out

    def buf(key, shape):
        a = out.get(key)
        if a is None:
            a = out[key] = _lib.pinned_empty(shape, np.float32)  # page-locked: full-rate D2H
        return a

    xf, Pf, Qf = buf("stateForward", (n, dim)), buf("stateCovarForward", (n, dim, dim)), buf("pNoiseForward", (n, dim, dim))
    D = buf("vectorD", (n,))
    xs, Ps = buf("stateSmoothed", (n, dim)), buf("stateCovarSmoothed", (n, dim, dim))
    lag = buf("lagCovSmoothed", (max(n - 1, 1), dim, dim))
    res = buf("postFitResiduals", (n, m)) if wantResiduals else None
    sum_d, sum_nll = C.c_double(0.0), C.c_double(0.0)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_sweep(
        ctx.handle, C.byref(mo), _ptr(data), _ptr(munc), m, n, _ptr(lam), _ptr(kap), _ptr(qs), _ptr(xf), _ptr(Pf),
        _ptr(Qf), _ptr(D), C.byref(sum_d), C.byref(sum_nll), _ptr(xs), _ptr(Ps), _ptr(lag), int(lag.shape[0]),
        _ptr(res)))
    out["phiHat"] = float(np.float32(sum_d.value / float(max(n, 1))))
    out["sumNLL"] = sum_nll.value
    return out


def _check_multiplier(init, n, what):
    """Validation half of the warm-start handling (cconsenrich.pyx:7899-7923)."""
    if init is None:
        return None
    src = np.asarray(init, dtype=np.float32).reshape(-1)
    if src.shape[0] != n:
        raise ValueError(f"{what} length must match intervalCount")
    if not np.all(np.isfinite(src)):
        raise ValueError(f"{what} must contain only finite values")
    return src


def _init_multiplier(src, n, lo, hi, pinned, fill=True):
    """Warm-start copy + clip (pyx:7899-7923).  The array travels to the device and back, so it is
    page-locked when the call is going to run on the device (``pinned``).  ``fill=False``: without a
    warm start the device sets the ones itself (cb200_ecm_opts.init_ones) and the array is output only."""
    arr = (_lib.pinned_empty if pinned else np.empty)((n,), np.float32)
    if src is None:
        if fill:
            arr.fill(1.0)
    else:
        np.clip(src, lo, hi, out=arr)
    return arr


def _ecm(dim, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount, stateInit,
         stateCovarInit, iters, rtol, pad, nu, lamMin, lamMax, kapMin, kapMax, useObs, useProc, useAPN,
         t_innerIters, returnIntermediates, returnDiagnostics, lambdaExpInit, processPrecExpInit,
         trackOptimizationPath, processQScale, apn=_APN_DEFAULTS):
    """ECM driver with the packing rules of cconsenrich.pyx:8409-8442."""
    data = _c32(matrixData, "matrixData", 2)
    munc = _c32(matrixPluginMuncInit, "matrixPluginMuncInit", 2)
    m, n = data.shape
    use_qscale = processQScale is not None
    lam = kap = None
    use_lam = bool(useObs)
    use_kap = bool(useProc and ((not useAPN) or use_qscale))
    lam_src = _check_multiplier(lambdaExpInit, n, "lambdaExpInit") if use_lam else None
    kap_src = _check_multiplier(processPrecExpInit, n, "processPrecExpInit") if use_kap else None
    qs = _coerce_qscale(processQScale, n) if use_qscale else None
    want = bool(returnIntermediates) and n > 0 and m > 0
    xs = Ps = lag = res = None  # allocated after validation (page-locked: the copy engine writes them)
    Q0 = np.asarray(matrixQ0)
    patience = 2
    path = [] if trackOptimizationPath else None

    def pack(iters_done, nll, diag):
        if returnIntermediates:
            out = (iters_done, float(nll), xs, Ps, lag, res, lam, kap)
            return out + (diag,) if returnDiagnostics else out
        return (iters_done, float(nll), diag) if returnDiagnostics else (iters_done, float(nll))

    if n <= 0 or m <= 0:
        lam = _init_multiplier(lam_src, n, lamMin, lamMax, False) if use_lam else None
        kap = _init_multiplier(kap_src, n, kapMin, kapMax, False) if use_kap else None
        xs, Ps = np.empty((n, dim), np.float32), np.empty((n, dim, dim), np.float32)
        lag, res = np.empty((max(n - 1, 1), dim, dim), np.float32), np.empty((n, m), np.float32)
        diag = {"iters_done": 0, "max_iters": int(iters), "converged": False, "skipped": True,
                "skip_reason": "too_few_intervals" if n > 0 else "empty_input", "fallback": "filter_smoother_only",
                "stable_iters": 0, "patience_target": patience, "initial_nll": 0.0, "final_nll": 0.0,
                "final_abs_rel_change": None, "final_rel_improvement": None, "nll_increase_count": 0}
        if trackOptimizationPath:
            diag["optimization_path"] = path
        return pack(0, 0.0, diag)

    # validation order of pyx:8131-8140 / 7396-7404
    if blockCount <= 0:
        raise ValueError("blockCount must be positive")
    if munc.shape[0] != m or munc.shape[1] != n:
        raise ValueError("matrixPluginMuncInit shape must match matrixData shape")
    if dim == 1 and float(Q0[0, 0]) <= 0.0:
        raise ValueError("matrixQ0[0, 0] must be positive")
    lamMin_d, lamMax_d, kapMin_d, kapMax_d = map(_f32, (lamMin, lamMax, kapMin, kapMax))
    _check_bounds(lamMin_d, lamMax_d, True)
    _check_bounds(kapMin_d, kapMax_d, False)
    bm = np.ascontiguousarray(intervalToBlockMap, dtype=np.int32)
    if bm.shape[0] < n:
        raise ValueError("intervalToBlockMap length must match intervalCount")
    if dim == 2:
        det = float(Q0[0, 0]) * float(Q0[1, 1]) - float(Q0[0, 1]) * float(Q0[1, 0])
        if det == 0.0:
            raise ValueError("matrixQ0 is singular")
    lam = _init_multiplier(lam_src, n, lamMin, lamMax, True, fill=False) if use_lam else None
    kap = _init_multiplier(kap_src, n, kapMin, kapMax, True, fill=False) if use_kap else None
    alloc = _lib.pinned_empty if want else np.empty
    xs, Ps = alloc((n, dim), np.float32), alloc((n, dim, dim), np.float32)
    lag, res = alloc((max(n - 1, 1), dim, dim), np.float32), alloc((n, m), np.float32)
    mo = _model(dim, matrixF, Q0, stateInit, stateCovarInit, pad, lamMin_d, lamMax_d, kapMin_d, kapMax_d,
                lam is not None, kap is not None, use_qscale, True, False, useAPN, apn)
    op = _lib.EcmOpts()
    op.max_iters, op.inner_iters = int(iters), int(t_innerIters)
    op.update_lambda, op.update_kappa = int(lam is not None), int(kap is not None)
    op.want_outputs = int(bool(returnIntermediates))
    op.init_ones = (1 if (lam is not None and lam_src is None) else 0) | (2 if (kap is not None and kap_src is None) else 0)
    op.rtol, op.nu = _f32(rtol), _f32(nu)
    result = _lib.EcmResult()
    nll_path = np.zeros(max(int(iters), 1), np.float64)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_ecm(
        ctx.handle, C.byref(mo), C.byref(op), _ptr(data), _ptr(munc), m, n, _ptr(bm), int(blockCount), _ptr(qs),
        _ptr(lam), _ptr(kap), _ptr(xs) if want else None, _ptr(Ps) if want else None, _ptr(lag) if want else None,
        _ptr(res) if want else None, C.byref(result), _ptr(nll_path)))

    if result.skipped:  # n <= 5 (pyx:7998-8129)
        diag = {"iters_done": 0, "max_iters": int(iters), "converged": False, "skipped": True,
                "skip_reason": "too_few_intervals", "fallback": "filter_smoother_only", "stable_iters": 0,
                "patience_target": patience, "initial_nll": float(result.final_nll),
                "final_nll": float(result.final_nll), "final_abs_rel_change": None, "final_rel_improvement": None,
                "nll_increase_count": 0}
        if trackOptimizationPath:
            diag["optimization_path"] = path
        return pack(0, result.final_nll, diag)

    if trackOptimizationPath:  # rebuilt from the per-iteration NLL values (pyx:8364-8392)
        prev, stable, rtol_d = None, 0, _f32(rtol)
        for i in range(result.iters_done):
            cur = float(nll_path[i])
            if prev is None:
                delta, scale = 0.0, max(abs(cur), 1.0)
            else:
                delta, scale = abs(cur - prev), max(abs(prev), abs(cur), 1.0)
            tol = rtol_d * scale
            stable = stable + 1 if (prev is not None and delta <= tol) else 0
            path.append({
                "iter": i + 1, "objective_name": "nll", "objective_value": cur,
                "change": float(delta) if prev is not None else None,
                "relative_improvement": float((prev - cur) / scale) if prev is not None else None,
                "abs_relative_change": float(delta / scale) if prev is not None else None,
                "threshold": float(tol) if prev is not None else None, "stable_iters": int(stable),
                "patience_target": patience, "reset_iteration": bool(prev is None),
                "converged": bool(stable >= patience),
            })
            prev = cur
    has_init = bool(result.has_initial)
    diag = {
        "iters_done": int(result.iters_done), "max_iters": int(iters), "converged": bool(result.converged),
        "skipped": False, "skip_reason": None, "fallback": None, "stable_iters": int(result.stable_iters),
        "patience_target": patience, "initial_nll": float(result.initial_nll) if has_init else None,
        "final_nll": float(result.final_nll),
        "final_abs_rel_change": float(result.final_abs_rel_change) if has_init else None,
        "final_rel_improvement": float(result.final_rel_improvement) if has_init else None,
        "nll_increase_count": int(result.nll_increase_count),
    }
    if trackOptimizationPath:
        diag["optimization_path"] = path
    return pack(int(result.iters_done), result.final_nll, diag)


def cfixedBackgroundECM(matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                        stateInit, stateCovarInit, ECM_fixedBackgroundIters=50, ECM_fixedBackgroundRtol=1.0e-4,
                        pad=1.0e-4, ECM_robustTNu=8.0, obsPrecisionMultiplierMin=0.25,
                        obsPrecisionMultiplierMax=4.0, procPrecisionMultiplierMin=0.25,
                        procPrecisionMultiplierMax=4.0, ECM_useObsPrecisionReweighting=True,
                        ECM_useProcessPrecisionReweighting=True, ECM_useAPN=False, APN_minQ=1.0e-4,
                        APN_maxQ=1000.0, APN_dStatThresh=5.0, APN_dStatScale=10.0, APN_dStatPC=2.0,
                        t_innerIters=5, returnIntermediates=False, returnDiagnostics=False,
                        lambdaExpInit=None, processPrecExpInit=None, trackOptimizationPath=False,
                        logIterations=True, processQScale=None):
    """2-state fixed-background ECM; signature and returns of cconsenrich.pyx:7660-8442."""
    return _ecm(2, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                stateInit, stateCovarInit, ECM_fixedBackgroundIters, ECM_fixedBackgroundRtol, pad, ECM_robustTNu,
                obsPrecisionMultiplierMin, obsPrecisionMultiplierMax, procPrecisionMultiplierMin,
                procPrecisionMultiplierMax, ECM_useObsPrecisionReweighting, ECM_useProcessPrecisionReweighting,
                ECM_useAPN, t_innerIters, returnIntermediates, returnDiagnostics, lambdaExpInit,
                processPrecExpInit, trackOptimizationPath, processQScale,
                (APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC))


def cfixedBackgroundECMLevel(matrixData, matrixPluginMuncInit, matrixQ0, intervalToBlockMap, blockCount,
                             stateInit, stateCovarInit, ECM_fixedBackgroundIters=50,
                             ECM_fixedBackgroundRtol=1.0e-4, pad=1.0e-4, ECM_robustTNu=8.0,
                             obsPrecisionMultiplierMin=0.25, obsPrecisionMultiplierMax=4.0,
                             procPrecisionMultiplierMin=0.25, procPrecisionMultiplierMax=4.0,
                             ECM_useObsPrecisionReweighting=True, ECM_useProcessPrecisionReweighting=True,
                             ECM_useAPN=False, APN_minQ=1.0e-4, APN_maxQ=1000.0, APN_dStatThresh=5.0,
                             APN_dStatScale=10.0, APN_dStatPC=2.0, t_innerIters=5, returnIntermediates=False,
                             returnDiagnostics=False, lambdaExpInit=None, processPrecExpInit=None,
                             trackOptimizationPath=False, logIterations=True, processQScale=None):
    """Level-only fixed-background ECM; signature and returns of cconsenrich.pyx:7153-7657."""
    return _ecm(1, matrixData, matrixPluginMuncInit, None, matrixQ0, intervalToBlockMap, blockCount,
                stateInit, stateCovarInit, ECM_fixedBackgroundIters, ECM_fixedBackgroundRtol, pad, ECM_robustTNu,
                obsPrecisionMultiplierMin, obsPrecisionMultiplierMax, procPrecisionMultiplierMin,
                procPrecisionMultiplierMax, ECM_useObsPrecisionReweighting, ECM_useProcessPrecisionReweighting,
                ECM_useAPN, t_innerIters, returnIntermediates, returnDiagnostics, lambdaExpInit,
                processPrecExpInit, trackOptimizationPath, processQScale,
                (APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC))


# ------------------------------------------------------------------------------------------
# background track: weighted statistics + roughness-penalised solve (core.py:8085-8378)
# ------------------------------------------------------------------------------------------
def _background_stats(residualMatrix, invVarMatrix, want_support):
    res = np.ascontiguousarray(residualMatrix, dtype=np.float32)
    inv = np.ascontiguousarray(invVarMatrix, dtype=np.float32)
    if res.ndim != 2 or inv.ndim != 2 or inv.shape[0] != res.shape[0] or inv.shape[1] != res.shape[1]:
        raise ValueError("residualMatrix and invVarMatrix must have identical 2D shapes")
    m, n = res.shape
    weight, rhs = np.empty(n, np.float64), np.empty(n, np.float64)
    support = C.c_int64(0)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_background_stats(ctx.handle, _ptr(res), _ptr(inv), m, n, _ptr(weight), _ptr(rhs),
                                                    C.byref(support) if want_support else None))
    return weight, rhs, int(support.value)


def cbackgroundWeightedStats(residualMatrix, invVarMatrix):
    """Column-wise background sufficient statistics; signature and returns of cconsenrich.pyx:9675-9697."""
    weight, rhs, _ = _background_stats(residualMatrix, invVarMatrix, False)
    return weight, rhs


def cbackgroundWeightedStatsWithSupport(residualMatrix, invVarMatrix):
    """As above plus the number of intervals with positive weight (cconsenrich.pyx:9700-9724)."""
    return _background_stats(residualMatrix, invVarMatrix, True)


def csolveZeroCenteredBackground(weightTrack, rhsTrack, lam, zeroCenter=True, lamFirst=0.0):
    """Roughness-penalised background update, ``(diag(w) + lamFirst D1'D1 + lam D2'D2) x = rhs`` with an
    optional zero-sum constraint; signature, checks and error texts of cconsenrich.pyx:944-1096."""
    w = np.ascontiguousarray(weightTrack, dtype=np.float64).reshape(-1)
    rhs = np.ascontiguousarray(rhsTrack, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    lam, lamFirst = float(lam), float(lamFirst)
    if rhs.shape[0] != n:
        raise ValueError("weightTrack and rhsTrack must have the same length")
    if not np.isfinite(lamFirst) or lamFirst < 0.0:
        raise ValueError("lamFirst must be finite and nonnegative")
    if not np.isfinite(lam) or lam < 0.0:
        raise ValueError("lam must be finite and nonnegative")
    out = np.zeros(n, np.float64)
    if n <= 0:
        return out
    bad, val = C.c_int64(-1), C.c_double(0.0)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_background_solve(ctx.handle, _ptr(w), _ptr(rhs), n, lam, lamFirst,
                                                    int(bool(zeroCenter)), _ptr(out), C.byref(bad), C.byref(val)))
    if bad.value >= 0:
        raise RuntimeError("roughness-penalized LDL factorization required pivot "
                           f"modification at index {bad.value} (pivot={val.value:.6g}, floor={1.0e-12:.6g}).")
    return out


# ------------------------------------------------------------------------------------------
# observation-noise (MUNC) stage: dense kernels (consenrich.py:7546)
# ------------------------------------------------------------------------------------------
def cMuncSmoothDenseLocalEvidence(localEvidence, windowIntervals, excludeMask=None, eps=1.0e-12):
    """Centred rolling mean of per-cell local evidence with an exclusion mask; signature, checks and
    error texts of cconsenrich.pyx:5642-5740."""
    local = np.asarray(localEvidence)
    if local.dtype != np.float32 or local.ndim != 2:
        raise ValueError("localEvidence must be a two-dimensional float32 array")
    local = np.ascontiguousarray(local)
    m, n = local.shape
    window = int(windowIntervals)
    eps_d = _f32(eps)  # the reference takes eps as a C float (pyx:5646)
    if window < 1:
        raise ValueError("windowIntervals must be positive")
    if eps_d <= 0.0 or not np.isfinite(eps_d):
        raise ValueError("eps must be positive and finite")
    mask, mode = None, 0
    if excludeMask is not None:
        mask = np.ascontiguousarray(excludeMask, dtype=np.uint8)
        if mask.ndim == 1:
            if mask.shape[0] != n:
                raise ValueError("excludeMask length must match interval count")
            mode = 1
        elif mask.ndim == 2:
            if mask.shape[0] != m or mask.shape[1] != n:
                raise ValueError("excludeMask shape must match localEvidence shape")
            mode = 2
        else:
            raise ValueError("excludeMask must be one- or two-dimensional")
    out = np.empty((m, n), np.float32)
    if m == 0 or n == 0:
        return out
    invalid = C.c_int32(0)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_munc_smooth_local_evidence(ctx.handle, _ptr(local), _ptr(mask), mode, m, n, window,
                                                              eps_d, _ptr(out), C.byref(invalid)))
    if invalid.value:
        raise ValueError("active local evidence cells must be positive and finite")
    return out


def _seed_pass_arguments(matrixData, matrixMunc, stateMean, stateVariance, background, gVariance, countFloor, omegaIn,
                         rhoIn, pad, studentTdf, useSeedWeights, updateWeights, omegaMin, omegaMax, varianceFloor,
                         varianceCap, enabled, studentT, dOmega, activeMask):
    """Argument handling of cMuncObservationMomentSeedPass (cconsenrich.pyx:5042-5198): coercions, checks in
    the reference's order with its error texts.  Returns the arrays (kept alive by the caller) and scalars."""
    data = np.asarray(matrixData)
    munc = np.asarray(matrixMunc)
    if data.dtype != np.float32 or data.ndim != 2 or munc.dtype != np.float32 or munc.ndim != 2:
        raise ValueError("matrixData and matrixMunc must be two-dimensional float32 arrays")
    data, munc = np.ascontiguousarray(data), np.ascontiguousarray(munc)
    mean, var = np.asarray(stateMean), np.asarray(stateVariance)
    if mean.dtype != np.float32 or mean.ndim != 1 or var.dtype != np.float32 or var.ndim != 1:
        raise ValueError("stateMean and stateVariance must be one-dimensional float32 arrays")
    mean, var = np.ascontiguousarray(mean), np.ascontiguousarray(var)
    m, n = data.shape
    use_weights = bool(enabled) and bool(useSeedWeights)
    student_t, update = bool(studentT), bool(updateWeights)
    pad_d, df, d_om = _f32(pad), _f32(studentTdf), _f32(dOmega)
    om_lo, om_hi, vfloor, vcap = _f32(omegaMin), _f32(omegaMax), _f32(varianceFloor), _f32(varianceCap)
    if munc.shape[0] != m or munc.shape[1] != n:
        raise ValueError("matrixMunc shape must match matrixData shape")
    if mean.shape[0] != n:
        raise ValueError("stateMean length must match interval count")
    if var.shape[0] != n:
        raise ValueError("stateVariance length must match interval count")
    if pad_d < 0.0 or not np.isfinite(pad_d):
        raise ValueError("pad must be finite and nonnegative")
    if vfloor <= 0.0 or not np.isfinite(vfloor):
        raise ValueError("varianceFloor must be positive and finite")
    if not np.isfinite(vcap) or vcap < vfloor:
        raise ValueError("varianceCap must be greater than or equal to varianceFloor")
    if use_weights and student_t and (df <= 0.0 or d_om <= 0.0 or not np.isfinite(df) or not np.isfinite(d_om)
                                      or om_lo <= 0.0 or om_hi < om_lo or not np.isfinite(om_lo)
                                      or not np.isfinite(om_hi)):
        raise ValueError("seed weight parameters are invalid")
    arrays = dict(data=data, munc=munc, state_mean=mean, state_var=var)
    if background is not None:
        arrays["background"] = np.ascontiguousarray(background, dtype=np.float32).reshape(-1)
        if arrays["background"].shape[0] != n:
            raise ValueError("background length must match interval count")
    if gVariance is not None:
        arrays["g_var"] = np.ascontiguousarray(gVariance, dtype=np.float32).reshape(-1)
        if arrays["g_var"].shape[0] != n:
            raise ValueError("gVariance length must match interval count")
    if countFloor is not None:
        cf = np.ascontiguousarray(countFloor, dtype=np.float32)
        if cf.ndim != 2 or cf.shape[0] != m or cf.shape[1] != n:
            raise ValueError("countFloor shape must match matrixData shape")
        arrays["count_floor"] = cf
    if omegaIn is not None:
        om = np.ascontiguousarray(omegaIn, dtype=np.float32)
        if om.ndim != 1:
            raise ValueError("omegaIn must be one-dimensional")
        if om.shape[0] != n:
            raise ValueError("omegaIn length must match interval count")
        arrays["omega_in"] = om
    if rhoIn is not None:
        rho = np.ascontiguousarray(rhoIn, dtype=np.float32)
        if rho.ndim != 2 or rho.shape[0] != m or rho.shape[1] != n:
            raise ValueError("rhoIn shape must match matrixData shape")
        arrays["rho_in"] = rho
    mode = 0
    if activeMask is not None:
        act = np.ascontiguousarray(activeMask, dtype=np.uint8)
        if act.ndim == 1:
            if act.shape[0] != n:
                raise ValueError("activeMask length must match interval count")
            mode = 1
        elif act.ndim == 2:
            if act.shape[0] != m or act.shape[1] != n:
                raise ValueError("activeMask shape must match matrixData shape")
            mode = 2
        else:
            raise ValueError("activeMask must be one- or two-dimensional")
        arrays["active"] = act
    scalars = dict(m=m, n=n, ld=n, active_ld=n, active_mode=mode, use_weights=int(use_weights), student_t=int(student_t),
                   update_weights=int(update), pad=pad_d, student_t_df=df, d_omega=d_om, omega_min=om_lo,
                   omega_max=om_hi, variance_floor=vfloor, variance_cap=vcap)
    return arrays, scalars


def cMuncObservationMomentSeedPass(matrixData, matrixMunc, stateMean, stateVariance, background=None, gVariance=None,
                                   countFloor=None, omegaIn=None, rhoIn=None, pad=1.0e-4, studentTdf=8.0,
                                   useSeedWeights=True, updateWeights=True, omegaMin=0.01, omegaMax=100.0,
                                   varianceFloor=1.0e-12, varianceCap=3.4028234663852886e38, enabled=True,
                                   studentT=True, dOmega=8.0, activeMask=None):
    """Moment / Student-t weight / local-variance pass of the MUNC seed smoother; signature, checks, error
    texts and the six returned arrays of cconsenrich.pyx:5042-5345."""
    arrays, scalars = _seed_pass_arguments(matrixData, matrixMunc, stateMean, stateVariance, background, gVariance,
                                           countFloor, omegaIn, rhoIn, pad, studentTdf, useSeedWeights, updateWeights,
                                           omegaMin, omegaMax, varianceFloor, varianceCap, enabled, studentT, dOmega,
                                           activeMask)
    m, n = scalars["m"], scalars["n"]
    big = _lib.pinned_empty if m * n * 4 >= (1 << 20) else np.empty
    outs = dict(moment=big((m, n), np.float32), rho_out=big((m, n), np.float32), omega_raw=np.empty(n, np.float32),
                omega_out=np.empty(n, np.float32), local=big((m, n), np.float32), variance=big((m, n), np.float32))
    if n > 0:
        args = _lib.MuncSeedArgs()
        for k, v in {**arrays, **outs}.items():
            setattr(args, k, v.ctypes.data if v.size else None)
        for k, v in scalars.items():
            setattr(args, k, v)
        invalid = C.c_int32(0)
        ctx = _ctx()
        _lib.check(ctx._lib.cb200_host_munc_seed_pass(ctx.handle, C.byref(args), C.byref(invalid)))
        if invalid.value:
            raise ValueError("active MUNC seed cells must be finite with positive denominators")
    return (outs["moment"], outs["rho_out"], outs["omega_raw"], outs["omega_out"], outs["local"], outs["variance"])


def cEMA(x, alpha):
    """Forward-then-backward exponential filter of a track; signature and dtype rule of
    cconsenrich.pyx:5897-5915 (a float32 ndarray stays float32, everything else is filtered as float64)."""
    is_f32 = isinstance(x, np.ndarray) and x.dtype == np.float32
    arr = np.ascontiguousarray(x, dtype=np.float32 if is_f32 else np.float64).reshape(-1)
    n = arr.shape[0]
    out = np.empty(n, arr.dtype)
    if n == 0:
        return out
    a = _f32(alpha) if is_f32 else float(alpha)
    if not (0.0 <= a <= 1.0):
        # the reference's kernel refuses such an alpha and its wrapper hands back the unwritten array
        # (pyx:5747-5748, 5908): there is no value to agree with, so say so instead
        raise ValueError("alpha must lie in [0, 1]")
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_ema(ctx.handle, _ptr(arr), n, int(not is_f32), float(alpha), _ptr(out)))
    return out


def cFinalizeMuncEBTrack(localVarianceTrack, priorVarianceTrack=None, countFloor=None, nuLocal=0.0, nuPrior=0.0,
                         varianceFloor=1.0e-12, varianceCap=3.4028234663852886e38, useEB=True):
    """Shrinkage of the local variance track towards its prior, clipping and count floor; signature,
    checks, error texts and diagnostics of cconsenrich.pyx:5445-5545."""
    local = np.ascontiguousarray(localVarianceTrack, dtype=np.float32).reshape(-1)
    n = local.shape[0]
    nu_l, nu_p = _f32(nuLocal), _f32(nuPrior)            # C float arguments (pyx:5449-5452)
    vfloor, vcap = _f32(varianceFloor), _f32(varianceCap)
    use_eb = bool(useEB)
    if vfloor <= 0.0 or not np.isfinite(vfloor):
        raise ValueError("varianceFloor must be positive and finite")
    if vcap < vfloor or not np.isfinite(vcap):
        raise ValueError("varianceCap must be finite and at least varianceFloor")
    prior = cfloor = None
    if use_eb:
        if priorVarianceTrack is None:
            raise ValueError("priorVarianceTrack is required for MUNC EB finalization")
        if not np.isfinite(nu_l) or nu_l <= 0.0:
            raise ValueError("nuLocal must be positive and finite")
        if not np.isfinite(nu_p) or nu_p <= 0.0:
            raise ValueError("nuPrior must be positive and finite")
        if not np.isfinite(nu_l + nu_p) or nu_l + nu_p <= 0.0:
            raise ValueError("posterior sample size must be positive and finite")
        prior = np.ascontiguousarray(priorVarianceTrack, dtype=np.float32).reshape(-1)
        if prior.shape[0] != n:
            raise ValueError("priorVarianceTrack length must match localVarianceTrack length")
    if countFloor is not None:
        cfloor = np.ascontiguousarray(countFloor, dtype=np.float32).reshape(-1)
        if cfloor.shape[0] != n:
            raise ValueError("countFloor length must match localVarianceTrack length")
    out = np.empty(n, np.float32)
    res = _lib.MuncFinalizeResult()
    res.invalid_local = res.invalid_prior = res.invalid_count_floor = -1
    if n > 0:
        ctx = _ctx()
        _lib.check(ctx._lib.cb200_host_munc_finalize_eb(ctx.handle, _ptr(local), _ptr(prior), _ptr(cfloor), n, nu_l, nu_p,
                                                        vfloor, vcap, int(use_eb), _ptr(out), C.byref(res)))
    if res.invalid_local >= 0:
        raise ValueError(f"localVarianceTrack must contain finite positive values at index {res.invalid_local}")
    if res.invalid_prior >= 0:
        raise ValueError(f"priorVarianceTrack must contain finite positive values at index {res.invalid_prior}")
    if res.invalid_count_floor >= 0:
        raise ValueError(f"countFloor must be nonnegative where finite at index {res.invalid_count_floor}")
    return out, {
        "supportCount": int(res.support_count),
        "supportFraction": (float(res.support_count) / float(n)) if n > 0 else 0.0,
        "countFloorFiniteCount": int(res.count_floor_finite),
        "countFloorAddedCount": int(res.count_floor_added),
        "countFloorMissingCount": int(res.count_floor_missing),
        "finalShrinkagePairCount": n if use_eb else 0,
        "finalShrinkagePairFraction": 1.0 if use_eb and n > 0 else 0.0,
    }


_saved: dict = {}


class HostPathWarning(RuntimeWarning):
    """An installed function was handed an input its device version does not cover and the reference's own
    host function (the one ``install`` replaced) ran instead.  Never silent."""


_MUNC_WINDOW_MAX = 8192  # csrc/munc_kernels.cu: tile + window must fit a CTA's shared memory


def _make_smooth_with_host_path(original):
    """cMuncSmoothDenseLocalEvidence for ``install``: windows beyond the device kernel's limit (a user-settable
    muncLocalWindowSizeBP far above the dependence spans the reference sizes it from) go to the function that
    was replaced, with a HostPathWarning, instead of failing a configuration the reference accepts."""
    import functools
    import warnings

    @functools.wraps(cMuncSmoothDenseLocalEvidence)
    def smooth(localEvidence, windowIntervals, excludeMask=None, eps=1.0e-12):
        if original is not None and int(windowIntervals) > _MUNC_WINDOW_MAX:
            warnings.warn(f"consenrich_b200: cMuncSmoothDenseLocalEvidence ran the reference's host implementation "
                          f"(windowIntervals {int(windowIntervals)} > {_MUNC_WINDOW_MAX})", HostPathWarning, stacklevel=2)
            return original(localEvidence, windowIntervals, excludeMask=excludeMask, eps=eps)
        return cMuncSmoothDenseLocalEvidence(localEvidence, windowIntervals, excludeMask=excludeMask, eps=eps)

    return smooth


def install(module=None, background=True, munc=True):
    """Replace the six hot-path attributes of ``consenrich.cconsenrich`` (or ``module``) with the
    B200 implementations.  ``core.py`` looks them up by attribute at call time (core.py:4274,
    4309, 3286), so ``runConsenrich`` picks them up without modification.  ``background``: also the
    three background-track functions (core.py:7543, 8145); ``munc``: also the dense kernels of the
    observation-noise stage (consenrich.py:7546)."""
    if module is None:
        import importlib
        module = importlib.import_module("consenrich.cconsenrich")
    _lib.load()  # fail now, loudly, if the native library is missing
    saved = _saved.setdefault(id(module), {})
    for name in _HOT_PATH + (_BACKGROUND if background else ()) + (_MUNC if munc else ()):
        if name not in saved:
            saved[name] = getattr(module, name, None)
        fn = globals()[name]
        if name == "cMuncSmoothDenseLocalEvidence":
            fn = _make_smooth_with_host_path(saved[name])
        setattr(module, name, fn)
    return module


def uninstall(module=None):
    if module is None:
        import importlib
        module = importlib.import_module("consenrich.cconsenrich")
    for name, fn in _saved.pop(id(module), {}).items():
        if fn is not None:
            setattr(module, name, fn)
    return module
