"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU oracle for the Consenrich state-space hot path: ctypes wrappers around
``oracle/ssm_oracle.c`` (a sequential C restatement of the loops in
``/root/reference/src/consenrich/cconsenrich.pyx``) exposing the six native entry points
with the reference's keyword signatures and return tuples:

* ``cforwardPass``            <- cconsenrich.pyx:6393-6632
* ``cbackwardPass``           <- cconsenrich.pyx:6635-6850
* ``cforwardPassLevel``       <- cconsenrich.pyx:6853-7049
* ``cbackwardPassLevel``      <- cconsenrich.pyx:7052-7150
* ``cfixedBackgroundECMLevel``<- cconsenrich.pyx:7153-7657
* ``cfixedBackgroundECM``     <- cconsenrich.pyx:7660-8442

and (``oracle/background_oracle.c``) the background-track functions on the other side of the ECM:

* ``cbackgroundWeightedStats[WithSupport]`` <- cconsenrich.pyx:9675-9724
* ``csolveZeroCenteredBackground``          <- cconsenrich.pyx:944-1096

and (``oracle/munc_oracle.c``) the dense kernels of the observation-noise stage:

* ``cMuncSmoothDenseLocalEvidence``         <- cconsenrich.pyx:5547-5740
* ``cFinalizeMuncEBTrack``                  <- cconsenrich.pyx:5365-5545
* ``cMuncObservationMomentSeedPass``        <- cconsenrich.pyx:4767-5345
* ``cEMA``                                  <- cconsenrich.pyx:5744-5759, 5897-5915

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product (``consenrich_b200``) never does.

Parity pinning: ``tests/test_oracle_pinning.py`` checks this restatement (a) bit-for-bit
against the reference itself (``oracle/_ref``, built by ``oracle/build_ref.sh`` from the
unmodified ``cconsenrich.pyx``) whenever that build is present, (b) against golden vectors
generated from that build (``tests/golden/``), and (c) against independent float64
known-answer recursions of the kind the reference's own tests use
(``tests/test_core.py:522-592``, tolerance 2e-6).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libssm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, reference flags).  Returns the library path."""
    srcs = [os.path.join(_HERE, f) for f in ("ssm_oracle.c", "background_oracle.c", "munc_oracle.c")]
    if force or (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("state_init", C.c_double),
        ("cov_init", C.c_double),
        ("pad", C.c_double),
        ("F", C.c_double * 4),
        ("Q0", C.c_double * 4),
        ("lam_min", C.c_double),
        ("lam_max", C.c_double),
        ("kap_min", C.c_double),
        ("kap_max", C.c_double),
        ("apn_min_q", C.c_double),
        ("apn_max_q", C.c_double),
        ("apn_thresh", C.c_double),
        ("apn_scale", C.c_double),
        ("apn_pc", C.c_double),
        ("use_lambda", C.c_int32),
        ("use_kappa", C.c_int32),
        ("use_qscale", C.c_int32),
        ("use_apn", C.c_int32),
        ("return_nll", C.c_int32),
        ("store_nll_in_d", C.c_int32),
        ("do_store", C.c_int32),
        ("pad_", C.c_int32),
    ]


_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        fp, ip, dp, vp = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p)
        for name in ("oracle_forward2", "oracle_forward1"):
            f = getattr(_lib, name)
            f.restype = C.c_int64
            f.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, vp,
                          C.POINTER(_Params), vp, vp, vp, vp, dp, dp]
        _lib.oracle_backward2.restype = None
        _lib.oracle_backward2.argtypes = [vp, C.c_int64, C.c_int64, dp, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
        _lib.oracle_backward1.restype = None
        _lib.oracle_backward1.argtypes = [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
        _lib.oracle_update_lambda.restype = None
        _lib.oracle_update_lambda.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int,
                                              C.c_double, C.c_double, C.c_double, C.c_double, vp]
        _lib.oracle_update_kappa2.restype = None
        _lib.oracle_update_kappa2.argtypes = [C.c_int64, vp, C.c_int64, vp, vp, vp, dp, dp, vp,
                                              C.c_double, C.c_double, C.c_double, vp]
        _lib.oracle_update_kappa1.restype = None
        _lib.oracle_update_kappa1.argtypes = [C.c_int64, vp, C.c_int64, vp, vp, vp, C.c_double, vp,
                                              C.c_double, C.c_double, C.c_double, vp]
        del fp, ip
    return _lib


def _f32(x):
    """Round a Python scalar to C float and widen back, as Cython does for ``float`` args."""
    return float(np.float32(x))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c32(a, name, ndim):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == ndim and a.flags.c_contiguous):
        raise ValueError(f"{name}: Buffer dtype mismatch / not C-contiguous float32 ndim={ndim}")
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
        if is_obs:
            raise ValueError("observation precision multiplier bounds must satisfy 0 < min <= max")
        raise ValueError("process precision multiplier bounds must satisfy 0 < min <= max")


def _forward(dim, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
             stateInit, stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD,
             returnNLL, storeNLLInD, lambdaExp, processPrecExp, useObs, useProc, useAPN,
             lamMin, lamMax, kapMin, kapMax, apnMinQ, apnMaxQ, apnThresh, apnScale, apnPC, processQScale):
    data = _c32(matrixData, "matrixData", 2)
    munc = _c32(matrixPluginMuncInit, "matrixPluginMuncInit", 2)
    m, n = data.shape
    do_store = stateForward is not None
    use_lambda = bool(useObs and (lambdaExp is not None))
    use_qscale = processQScale is not None
    use_kappa = bool(useProc and (processPrecExp is not None) and ((not useAPN) or use_qscale))
    qs = _coerce_qscale(processQScale, n) if use_qscale else None
    if n <= 0 or m <= 0:
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
    bm = intervalToBlockMap
    if bm.shape[0] < n:
        raise ValueError("intervalToBlockMap length must match intervalCount")
    if use_lambda and lambdaExp.shape[0] != n:
        raise ValueError("lambdaExp length must match intervalCount")
    if use_kappa and processPrecExp.shape[0] != n:
        raise ValueError("processPrecExp length must match intervalCount")
    if vectorD is None:
        d = np.empty(n, dtype=np.float32)
        vectorD = d
    else:
        d = vectorD
        if d.shape[0] < n:
            raise ValueError("vectorD length must match intervalCount")
    p = _Params()
    p.state_init, p.cov_init, p.pad = _f32(stateInit), _f32(stateCovarInit), _f32(pad)
    if dim == 2:
        p.F[0], p.F[1], p.F[2], p.F[3] = (float(Fm[0, 0]), float(Fm[0, 1]), float(Fm[1, 0]), float(Fm[1, 1]))
        p.Q0[0], p.Q0[1], p.Q0[2], p.Q0[3] = (float(Q0[0, 0]), float(Q0[0, 1]), float(Q0[1, 0]), float(Q0[1, 1]))
    else:
        p.Q0[0] = float(Q0[0, 0])
    p.lam_min, p.lam_max, p.kap_min, p.kap_max = lamMin, lamMax, kapMin, kapMax
    p.apn_min_q, p.apn_max_q, p.apn_thresh, p.apn_scale, p.apn_pc = map(
        _f32, (apnMinQ, apnMaxQ, apnThresh, apnScale, apnPC))
    p.use_lambda, p.use_kappa, p.use_qscale, p.use_apn = int(use_lambda), int(use_kappa), int(use_qscale), int(bool(useAPN))
    p.return_nll, p.store_nll_in_d, p.do_store = int(bool(returnNLL)), int(bool(storeNLLInD)), int(do_store)
    sum_d, sum_nll = C.c_double(0.0), C.c_double(0.0)
    fn = _L().oracle_forward2 if dim == 2 else _L().oracle_forward1
    bad = fn(_ptr(data), _ptr(munc), m, n, _ptr(bm), int(blockCount),
             _ptr(lambdaExp) if use_lambda else None, _ptr(processPrecExp) if use_kappa else None,
             _ptr(qs), C.byref(p), _ptr(d),
             _ptr(stateForward) if do_store else None, _ptr(stateCovarForward) if do_store else None,
             _ptr(pNoiseForward) if do_store else None, C.byref(sum_d), C.byref(sum_nll))
    if bad >= 0:
        raise ValueError("intervalToBlockMap has out-of-range block id")
    phi = np.float32(sum_d.value / float(n))
    phi = float(phi)
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
    return _forward(2, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                    stateInit, stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD,
                    returnNLL, storeNLLInD, lambdaExp, processPrecExp, ECM_useObsPrecisionReweighting,
                    ECM_useProcessPrecisionReweighting, ECM_useAPN, obsPrecisionMultiplierMin,
                    obsPrecisionMultiplierMax, procPrecisionMultiplierMin, procPrecisionMultiplierMax,
                    APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC, processQScale)


def cforwardPassLevel(matrixData, matrixPluginMuncInit, matrixQ0, intervalToBlockMap, blockCount,
                      stateInit, stateCovarInit, pad=1.0e-4, chunkSize=1000000, stateForward=None,
                      stateCovarForward=None, pNoiseForward=None, vectorD=None, returnNLL=False,
                      storeNLLInD=False, lambdaExp=None, processPrecExp=None,
                      ECM_useObsPrecisionReweighting=True, ECM_useProcessPrecisionReweighting=True,
                      ECM_useAPN=False, obsPrecisionMultiplierMin=0.25, obsPrecisionMultiplierMax=4.0,
                      procPrecisionMultiplierMin=0.25, procPrecisionMultiplierMax=4.0, APN_minQ=1.0e-4,
                      APN_maxQ=1000.0, APN_dStatThresh=5.0, APN_dStatScale=10.0, APN_dStatPC=2.0,
                      processQScale=None):
    return _forward(1, matrixData, matrixPluginMuncInit, None, matrixQ0, intervalToBlockMap, blockCount,
                    stateInit, stateCovarInit, pad, stateForward, stateCovarForward, pNoiseForward, vectorD,
                    returnNLL, storeNLLInD, lambdaExp, processPrecExp, ECM_useObsPrecisionReweighting,
                    ECM_useProcessPrecisionReweighting, ECM_useAPN, obsPrecisionMultiplierMin,
                    obsPrecisionMultiplierMax, procPrecisionMultiplierMin, procPrecisionMultiplierMax,
                    APN_minQ, APN_maxQ, APN_dStatThresh, APN_dStatScale, APN_dStatPC, processQScale)


def _backward(dim, matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward,
              stateSmoothed, stateCovarSmoothed, lagCovSmoothed, postFitResiduals):
    data = _c32(matrixData, "matrixData", 2)
    m, n = data.shape
    xs = stateSmoothed if stateSmoothed is not None else np.empty((n, dim), dtype=np.float32)
    Ps = stateCovarSmoothed if stateCovarSmoothed is not None else np.empty((n, dim, dim), dtype=np.float32)
    lag = lagCovSmoothed if lagCovSmoothed is not None else np.empty((max(n - 1, 1), dim, dim), dtype=np.float32)
    res = postFitResiduals if postFitResiduals is not None else np.empty((n, m), dtype=np.float32)
    if n <= 0:
        return (xs, Ps, lag, res)
    if dim == 2:
        Fm = np.asarray(matrixF)
        F = (C.c_double * 4)(float(Fm[0, 0]), float(Fm[0, 1]), float(Fm[1, 0]), float(Fm[1, 1]))
        _L().oracle_backward2(_ptr(data), m, n, F, _ptr(stateForward), _ptr(stateCovarForward),
                              _ptr(pNoiseForward), _ptr(xs), _ptr(Ps), _ptr(lag), int(lag.shape[0]), _ptr(res))
    else:
        _L().oracle_backward1(_ptr(data), m, n, _ptr(stateForward), _ptr(stateCovarForward),
                              _ptr(pNoiseForward), _ptr(xs), _ptr(Ps), _ptr(lag), int(lag.shape[0]), _ptr(res))
    return (xs, Ps, lag, res)


def cbackwardPass(matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward, chunkSize=1000000,
                  stateSmoothed=None, stateCovarSmoothed=None, lagCovSmoothed=None, postFitResiduals=None):
    return _backward(2, matrixData, matrixF, stateForward, stateCovarForward, pNoiseForward,
                     stateSmoothed, stateCovarSmoothed, lagCovSmoothed, postFitResiduals)


def cbackwardPassLevel(matrixData, stateForward, stateCovarForward, pNoiseForward, chunkSize=1000000,
                       stateSmoothed=None, stateCovarSmoothed=None, lagCovSmoothed=None,
                       postFitResiduals=None):
    return _backward(1, matrixData, None, stateForward, stateCovarForward, pNoiseForward,
                     stateSmoothed, stateCovarSmoothed, lagCovSmoothed, postFitResiduals)


def _init_multiplier(init, n, lo, hi, what):
    """Warm-start copy + clip, cconsenrich.pyx:7899-7923."""
    if init is None:
        return np.ones(n, dtype=np.float32)
    arr = np.array(init, dtype=np.float32, copy=True, order="C").reshape(-1)
    if arr.shape[0] != n:
        raise ValueError(f"{what} length must match intervalCount")
    if not np.all(np.isfinite(arr)):
        raise ValueError(f"{what} must contain only finite values")
    np.clip(arr, lo, hi, out=arr)
    return arr


def _ecm(dim, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
         stateInit, stateCovarInit, iters, rtol, pad, nu, lamMin, lamMax, kapMin, kapMax, useObs, useProc,
         useAPN, apnMinQ, apnMaxQ, apnThresh, apnScale, apnPC, t_innerIters, returnIntermediates,
         returnDiagnostics, lambdaExpInit, processPrecExpInit, trackOptimizationPath, processQScale):
    """ECM driver, cconsenrich.pyx:7877-8442 (2-state) / 7188-7657 (level)."""
    data = _c32(matrixData, "matrixData", 2)
    munc = _c32(matrixPluginMuncInit, "matrixPluginMuncInit", 2)
    m, n = data.shape
    use_qscale = processQScale is not None
    lam = kap = None
    if useObs:
        lam = _init_multiplier(lambdaExpInit, n, lamMin, lamMax, "lambdaExpInit")
    if useProc and ((not useAPN) or use_qscale):
        kap = _init_multiplier(processPrecExpInit, n, kapMin, kapMax, "processPrecExpInit")
    qs = _coerce_qscale(processQScale, n) if use_qscale else None
    xf = np.empty((n, dim), np.float32)
    Pf = np.empty((n, dim, dim), np.float32)
    Qf = np.empty((n, dim, dim), np.float32)
    xs = np.empty((n, dim), np.float32)
    Ps = np.empty((n, dim, dim), np.float32)
    lag = np.empty((max(n - 1, 1), dim, dim), np.float32)
    res = np.empty((n, m), np.float32)
    Q0 = np.asarray(matrixQ0)
    rtol_d, pad_d, nu_d = _f32(rtol), _f32(pad), _f32(nu)
    lamMin_d, lamMax_d, kapMin_d, kapMax_d = map(_f32, (lamMin, lamMax, kapMin, kapMax))
    patience = 2
    path = [] if trackOptimizationPath else None

    fkw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=matrixQ0,
               intervalToBlockMap=intervalToBlockMap, blockCount=blockCount, stateInit=stateInit,
               stateCovarInit=stateCovarInit, pad=pad, chunkSize=0, lambdaExp=lam, processPrecExp=kap,
               ECM_useObsPrecisionReweighting=useObs, ECM_useProcessPrecisionReweighting=useProc,
               ECM_useAPN=useAPN, obsPrecisionMultiplierMin=lamMin, obsPrecisionMultiplierMax=lamMax,
               procPrecisionMultiplierMin=kapMin, procPrecisionMultiplierMax=kapMax, APN_minQ=apnMinQ,
               APN_maxQ=apnMaxQ, APN_dStatThresh=apnThresh, APN_dStatScale=apnScale, APN_dStatPC=apnPC,
               processQScale=qs)
    if dim == 2:
        fkw["matrixF"] = matrixF
        fwd, bwd = cforwardPass, cbackwardPass
        bkw = dict(matrixData=data, matrixF=matrixF)
    else:
        fwd, bwd = cforwardPassLevel, cbackwardPassLevel
        bkw = dict(matrixData=data)

    def sweep():
        nonlocal xs, Ps, lag, res
        fwd(**fkw, stateForward=xf, stateCovarForward=Pf, pNoiseForward=Qf, vectorD=None,
            returnNLL=False, storeNLLInD=False)
        xs, Ps, lag, res = bwd(**bkw, stateForward=xf, stateCovarForward=Pf, pNoiseForward=Qf, chunkSize=0,
                               stateSmoothed=xs, stateCovarSmoothed=Ps, lagCovSmoothed=lag,
                               postFitResiduals=res)

    def nll_only():
        return float(fwd(**fkw, stateForward=None, stateCovarForward=None, pNoiseForward=None,
                         vectorD=None, returnNLL=True, storeNLLInD=False)[3])

    def validate():
        if blockCount <= 0:
            raise ValueError("blockCount must be positive")
        if dim == 1:
            if munc.shape[0] != m or munc.shape[1] != n:
                raise ValueError("matrixPluginMuncInit shape must match matrixData shape")
            if float(Q0[0, 0]) <= 0.0:
                raise ValueError("matrixQ0[0, 0] must be positive")
        _check_bounds(lamMin_d, lamMax_d, True)
        _check_bounds(kapMin_d, kapMax_d, False)
        if intervalToBlockMap.shape[0] < n:
            raise ValueError("intervalToBlockMap length must match intervalCount")
        if dim == 2:
            if munc.shape[0] != m or munc.shape[1] != n:
                raise ValueError("matrixPluginMuncInit shape must match matrixData shape")
            det = float(Q0[0, 0]) * float(Q0[1, 1]) - float(Q0[0, 1]) * float(Q0[1, 0])
            if det == 0.0:
                raise ValueError("matrixQ0 is singular")

    def pack(iters_done, nll, diag):
        if returnIntermediates:
            out = (iters_done, float(nll), xs, Ps, lag, res, lam, kap)
            return out + (diag,) if returnDiagnostics else out
        return (iters_done, float(nll), diag) if returnDiagnostics else (iters_done, float(nll))

    if n <= 5:  # pyx:7998-8129: filter + smoother only
        cur = 0.0
        if n > 0 and m > 0:
            validate()
            sweep()
            cur = nll_only()
        diag = {
            "iters_done": 0, "max_iters": int(iters), "converged": False, "skipped": True,
            "skip_reason": "too_few_intervals" if n > 0 else "empty_input",
            "fallback": "filter_smoother_only", "stable_iters": 0, "patience_target": patience,
            "initial_nll": float(cur), "final_nll": float(cur), "final_abs_rel_change": None,
            "final_rel_improvement": None, "nll_increase_count": 0,
        }
        if trackOptimizationPath:
            diag["optimization_path"] = path
        return pack(0, cur, diag)

    validate()
    Fd = None
    if dim == 2:
        Fm = np.asarray(matrixF)
        Fd = (C.c_double * 4)(float(Fm[0, 0]), float(Fm[0, 1]), float(Fm[1, 0]), float(Fm[1, 1]))
        Qd = (C.c_double * 4)(float(Q0[0, 0]), float(Q0[0, 1]), float(Q0[1, 0]), float(Q0[1, 1]))
    prev, cur, init_nll = 1.0e16, 0.0, 0.0
    has_init = False
    iters_done = stable = inc = 0
    converged = False
    rel_impr = abs_rel = 0.0
    L = _L()
    for i in range(int(iters)):
        iters_done = i + 1
        for _ in range(int(t_innerIters)):
            sweep()
            if useObs:
                L.oracle_update_lambda(_ptr(data), _ptr(munc), m, n, _ptr(intervalToBlockMap), int(blockCount),
                                       _ptr(xs), _ptr(Ps), dim, pad_d, nu_d, lamMin_d, lamMax_d, _ptr(lam))
            if kap is not None:
                if dim == 2:
                    L.oracle_update_kappa2(n, _ptr(intervalToBlockMap), int(blockCount), _ptr(xs), _ptr(Ps),
                                           _ptr(lag), Fd, Qd, _ptr(qs), nu_d, kapMin_d, kapMax_d, _ptr(kap))
                else:
                    L.oracle_update_kappa1(n, _ptr(intervalToBlockMap), int(blockCount), _ptr(xs), _ptr(Ps),
                                           _ptr(lag), float(Q0[0, 0]), _ptr(qs), nu_d, kapMin_d, kapMax_d,
                                           _ptr(kap))
        cur = nll_only()
        has_prev = has_init
        if not has_prev:
            init_nll, has_init = cur, True
        elif cur > prev + (1.0e-12 * max(abs(prev), 1.0)):
            inc += 1
        if has_prev:
            delta, scale = abs(cur - prev), abs(prev)
        else:
            delta, scale = 0.0, abs(cur)
        scale = max(scale, abs(cur))
        scale = max(scale, 1.0)
        if has_prev:
            rel_impr, abs_rel = (prev - cur) / scale, delta / scale
        else:
            rel_impr = abs_rel = 0.0
        tol = rtol_d * scale
        prev = cur
        stable = stable + 1 if (has_prev and delta <= tol) else 0
        it_conv = stable >= patience
        if trackOptimizationPath:
            path.append({
                "iter": iters_done, "objective_name": "nll", "objective_value": float(cur),
                "change": float(delta) if has_prev else None,
                "relative_improvement": float(rel_impr) if has_prev else None,
                "abs_relative_change": float(abs_rel) if has_prev else None,
                "threshold": float(tol) if has_prev else None, "stable_iters": int(stable),
                "patience_target": patience, "reset_iteration": bool(not has_prev),
                "converged": bool(it_conv),
            })
        if it_conv:
            converged = True
            break
    diag = {
        "iters_done": int(iters_done), "max_iters": int(iters), "converged": bool(converged),
        "skipped": False, "skip_reason": None, "fallback": None, "stable_iters": int(stable),
        "patience_target": patience, "initial_nll": float(init_nll) if has_init else None,
        "final_nll": float(prev), "final_abs_rel_change": float(abs_rel) if has_init else None,
        "final_rel_improvement": float(rel_impr) if has_init else None, "nll_increase_count": int(inc),
    }
    if trackOptimizationPath:
        diag["optimization_path"] = path
    return pack(iters_done, prev, diag)


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
    return _ecm(2, matrixData, matrixPluginMuncInit, matrixF, matrixQ0, intervalToBlockMap, blockCount,
                stateInit, stateCovarInit, ECM_fixedBackgroundIters, ECM_fixedBackgroundRtol, pad,
                ECM_robustTNu, obsPrecisionMultiplierMin, obsPrecisionMultiplierMax,
                procPrecisionMultiplierMin, procPrecisionMultiplierMax, ECM_useObsPrecisionReweighting,
                ECM_useProcessPrecisionReweighting, ECM_useAPN, APN_minQ, APN_maxQ, APN_dStatThresh,
                APN_dStatScale, APN_dStatPC, t_innerIters, returnIntermediates, returnDiagnostics,
                lambdaExpInit, processPrecExpInit, trackOptimizationPath, processQScale)


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
    return _ecm(1, matrixData, matrixPluginMuncInit, None, matrixQ0, intervalToBlockMap, blockCount,
                stateInit, stateCovarInit, ECM_fixedBackgroundIters, ECM_fixedBackgroundRtol, pad,
                ECM_robustTNu, obsPrecisionMultiplierMin, obsPrecisionMultiplierMax,
                procPrecisionMultiplierMin, procPrecisionMultiplierMax, ECM_useObsPrecisionReweighting,
                ECM_useProcessPrecisionReweighting, ECM_useAPN, APN_minQ, APN_maxQ, APN_dStatThresh,
                APN_dStatScale, APN_dStatPC, t_innerIters, returnIntermediates, returnDiagnostics,
                lambdaExpInit, processPrecExpInit, trackOptimizationPath, processQScale)


def load_reference():
    """Return the reference-built ``cconsenrich`` module from oracle/_ref, or None if absent."""
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "consenrich_ref")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import consenrich_ref.cconsenrich as ref  # type: ignore
    except Exception:
        return None
    return ref


# ------------------------------------------------------------------------------------------
# background track (oracle/background_oracle.c)
# ------------------------------------------------------------------------------------------
def _bg_lib():
    lib = _L()
    if not getattr(lib, "_bg_ready", False):
        lib.bg_weighted_stats.restype = C.c_int64
        lib.bg_weighted_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.bg_solve.restype = C.c_int64
        lib.bg_solve.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        lib._bg_ready = True
    return lib


def cbackgroundWeightedStatsWithSupport(residualMatrix, invVarMatrix):
    """cconsenrich.pyx:9700-9724."""
    res = np.ascontiguousarray(residualMatrix, dtype=np.float32)
    inv = np.ascontiguousarray(invVarMatrix, dtype=np.float32)
    if res.ndim != 2 or inv.ndim != 2 or inv.shape != res.shape:
        raise ValueError("residualMatrix and invVarMatrix must have identical 2D shapes")
    m, n = res.shape
    weight, rhs = np.empty(n, np.float64), np.empty(n, np.float64)
    support = _bg_lib().bg_weighted_stats(res.ctypes.data, inv.ctypes.data, m, n, weight.ctypes.data, rhs.ctypes.data)
    return weight, rhs, int(support)


def cbackgroundWeightedStats(residualMatrix, invVarMatrix):
    """cconsenrich.pyx:9675-9697."""
    return cbackgroundWeightedStatsWithSupport(residualMatrix, invVarMatrix)[:2]


def csolveZeroCenteredBackground(weightTrack, rhsTrack, lam, zeroCenter=True, lamFirst=0.0):
    """cconsenrich.pyx:944-1096 (checks and error texts included)."""
    w = np.ascontiguousarray(weightTrack, dtype=np.float64).reshape(-1)
    r = np.ascontiguousarray(rhsTrack, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    lam, lamFirst = float(lam), float(lamFirst)
    min_pivot = 1.0e-12
    if r.shape[0] != n:
        raise ValueError("weightTrack and rhsTrack must have the same length")
    if not np.isfinite(lamFirst) or lamFirst < 0.0:
        raise ValueError("lamFirst must be finite and nonnegative")
    if not np.isfinite(lam) or lam < 0.0:
        raise ValueError("lam must be finite and nonnegative")
    out = np.zeros(n, np.float64)
    if n <= 0:
        return out
    bad, val = -1, 0.0
    if n == 1:
        if not zeroCenter:
            if w[0] < min_pivot:
                bad, val = 0, float(w[0])
            else:
                out[0] = r[0] / w[0]
    else:
        diag, rhs = w.copy(), r.copy()
        cons, lower = np.empty(n, np.float64), np.empty(n, np.float64)
        v = C.c_double(0.0)
        bad = int(_bg_lib().bg_solve(diag.ctypes.data, rhs.ctypes.data, cons.ctypes.data, lower.ctypes.data,
                                     out.ctypes.data, n, lam, lamFirst, int(bool(zeroCenter)), C.byref(v)))
        val = v.value
    if bad >= 0:
        raise RuntimeError("roughness-penalized LDL factorization required pivot "
                           f"modification at index {bad} (pivot={val:.6g}, floor={min_pivot:.6g}).")
    return out


# ------------------------------------------------------------------------------------------
# observation-noise (MUNC) stage (oracle/munc_oracle.c)
# ------------------------------------------------------------------------------------------
def cMuncSmoothDenseLocalEvidence(localEvidence, windowIntervals, excludeMask=None, eps=1.0e-12):
    """cconsenrich.pyx:5642-5740 (checks and error texts included)."""
    local = np.asarray(localEvidence)
    if local.dtype != np.float32 or local.ndim != 2:
        raise ValueError("localEvidence must be a two-dimensional float32 array")
    local = np.ascontiguousarray(local)
    m, n = local.shape
    window = int(windowIntervals)
    eps_d = float(np.float32(eps))
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
    lib = _L()
    lib.munc_smooth_invalid_index.restype = C.c_int64
    lib.munc_smooth_invalid_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64]
    lib.munc_smooth_rows.restype = None
    lib.munc_smooth_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                     C.c_void_p]
    mp = mask.ctypes.data if mask is not None else None
    if lib.munc_smooth_invalid_index(local.ctypes.data, mp, mode, m, n) >= 0:
        raise ValueError("active local evidence cells must be positive and finite")
    out = np.empty((m, n), np.float32)
    lib.munc_smooth_rows(local.ctypes.data, mp, mode, m, n, window, eps_d, out.ctypes.data)
    return out


def cFinalizeMuncEBTrack(localVarianceTrack, priorVarianceTrack=None, countFloor=None, nuLocal=0.0, nuPrior=0.0,
                         varianceFloor=1.0e-12, varianceCap=3.4028234663852886e38, useEB=True):
    """cconsenrich.pyx:5445-5545 (checks, error texts and diagnostics included)."""
    f32 = lambda v: float(np.float32(v))
    local = np.ascontiguousarray(localVarianceTrack, dtype=np.float32).reshape(-1)
    n = local.shape[0]
    nu_l, nu_p, vfloor, vcap = f32(nuLocal), f32(nuPrior), f32(varianceFloor), f32(varianceCap)
    post = nu_l + nu_p
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
        if not np.isfinite(post) or post <= 0.0:
            raise ValueError("posterior sample size must be positive and finite")
        prior = np.ascontiguousarray(priorVarianceTrack, dtype=np.float32).reshape(-1)
        if prior.shape[0] != n:
            raise ValueError("priorVarianceTrack length must match localVarianceTrack length")
    if countFloor is not None:
        cfloor = np.ascontiguousarray(countFloor, dtype=np.float32).reshape(-1)
        if cfloor.shape[0] != n:
            raise ValueError("countFloor length must match localVarianceTrack length")
    out = np.empty(n, np.float32)
    counters, invalid = np.zeros(4, np.int64), np.full(3, -1, np.int64)
    lib = _L()
    lib.munc_finalize_eb.restype = None
    lib.munc_finalize_eb.argtypes = [C.c_void_p] * 4 + [C.c_int64] + [C.c_double] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
    lib.munc_finalize_eb(local.ctypes.data, prior.ctypes.data if prior is not None else None,
                         cfloor.ctypes.data if cfloor is not None else None, out.ctypes.data, n, nu_l, nu_p, post,
                         vfloor, vcap, int(use_eb), counters.ctypes.data, invalid.ctypes.data)
    if invalid[0] >= 0:
        raise ValueError(f"localVarianceTrack must contain finite positive values at index {int(invalid[0])}")
    if invalid[1] >= 0:
        raise ValueError(f"priorVarianceTrack must contain finite positive values at index {int(invalid[1])}")
    if invalid[2] >= 0:
        raise ValueError(f"countFloor must be nonnegative where finite at index {int(invalid[2])}")
    return out, {
        "supportCount": int(counters[0]),
        "supportFraction": (float(counters[0]) / float(n)) if n > 0 else 0.0,
        "countFloorFiniteCount": int(counters[1]),
        "countFloorAddedCount": int(counters[2]),
        "countFloorMissingCount": int(counters[3]),
        "finalShrinkagePairCount": n if use_eb else 0,
        "finalShrinkagePairFraction": 1.0 if use_eb and n > 0 else 0.0,
    }


class _SeedArgs(C.Structure):
    _fields_ = ([(k, C.c_void_p) for k in ("data", "munc", "state_mean", "state_var", "background", "g_var",
                                           "count_floor", "omega_in", "rho_in", "active", "moment", "rho_out",
                                           "omega_raw", "omega_out", "local", "variance")]
                + [("m", C.c_int64), ("n", C.c_int64)]
                + [(k, C.c_int32) for k in ("active_mode", "use_weights", "student_t", "update_weights")]
                + [(k, C.c_double) for k in ("pad", "d_s", "d_omega", "omega_min", "omega_max", "var_floor", "var_cap")])


def cMuncObservationMomentSeedPass(matrixData, matrixMunc, stateMean, stateVariance, background=None, gVariance=None,
                                   countFloor=None, omegaIn=None, rhoIn=None, pad=1.0e-4, studentTdf=8.0,
                                   useSeedWeights=True, updateWeights=True, omegaMin=0.01, omegaMax=100.0,
                                   varianceFloor=1.0e-12, varianceCap=3.4028234663852886e38, enabled=True,
                                   studentT=True, dOmega=8.0, activeMask=None):
    """cconsenrich.pyx:5042-5345 (checks and error texts included)."""
    f32 = lambda v: float(np.float32(v))
    data = np.ascontiguousarray(matrixData, dtype=np.float32)
    munc = np.ascontiguousarray(matrixMunc, dtype=np.float32)
    mean = np.ascontiguousarray(stateMean, dtype=np.float32)
    var = np.ascontiguousarray(stateVariance, dtype=np.float32)
    m, n = data.shape
    use_weights, student_t, update = bool(enabled) and bool(useSeedWeights), bool(studentT), bool(updateWeights)
    pad_d, df, d_om = f32(pad), f32(studentTdf), f32(dOmega)
    om_lo, om_hi, vfloor, vcap = f32(omegaMin), f32(omegaMax), f32(varianceFloor), f32(varianceCap)
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
    keep = dict(data=data, munc=munc, state_mean=mean, state_var=var)
    if background is not None:
        keep["background"] = np.ascontiguousarray(background, dtype=np.float32).reshape(-1)
        if keep["background"].shape[0] != n:
            raise ValueError("background length must match interval count")
    if gVariance is not None:
        keep["g_var"] = np.ascontiguousarray(gVariance, dtype=np.float32).reshape(-1)
        if keep["g_var"].shape[0] != n:
            raise ValueError("gVariance length must match interval count")
    if countFloor is not None:
        cf = np.ascontiguousarray(countFloor, dtype=np.float32)
        if cf.ndim != 2 or cf.shape[0] != m or cf.shape[1] != n:
            raise ValueError("countFloor shape must match matrixData shape")
        keep["count_floor"] = cf
    if omegaIn is not None:
        om = np.ascontiguousarray(omegaIn, dtype=np.float32)
        if om.ndim != 1:
            raise ValueError("omegaIn must be one-dimensional")
        if om.shape[0] != n:
            raise ValueError("omegaIn length must match interval count")
        keep["omega_in"] = om
    if rhoIn is None and use_weights and student_t and not update:
        keep["rho_in"] = np.ones((m, n), np.float32)
    elif rhoIn is not None:
        rho = np.ascontiguousarray(rhoIn, dtype=np.float32)
        if rho.ndim != 2 or rho.shape[0] != m or rho.shape[1] != n:
            raise ValueError("rhoIn shape must match matrixData shape")
        keep["rho_in"] = rho
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
        keep["active"] = act
    outs = dict(moment=np.empty((m, n), np.float32), rho_out=np.empty((m, n), np.float32),
                omega_raw=np.empty(n, np.float32), omega_out=np.empty(n, np.float32),
                local=np.empty((m, n), np.float32), variance=np.empty((m, n), np.float32))
    a = _SeedArgs()
    for k, v in {**keep, **outs}.items():
        setattr(a, k, v.ctypes.data)
    a.m, a.n, a.active_mode = m, n, mode
    a.use_weights, a.student_t, a.update_weights = int(use_weights), int(student_t), int(update)
    a.pad, a.d_s, a.d_omega, a.omega_min, a.omega_max, a.var_floor, a.var_cap = pad_d, df, d_om, om_lo, om_hi, vfloor, vcap
    lib = _L()
    lib.munc_seed_invalid_index.restype = C.c_int64
    lib.munc_seed_invalid_index.argtypes = [C.POINTER(_SeedArgs)]
    lib.munc_seed_pass.restype = None
    lib.munc_seed_pass.argtypes = [C.POINTER(_SeedArgs)]
    if lib.munc_seed_invalid_index(C.byref(a)) >= 0:
        raise ValueError("active MUNC seed cells must be finite with positive denominators")
    lib.munc_seed_pass(C.byref(a))
    return (outs["moment"], outs["rho_out"], outs["omega_raw"], outs["omega_out"], outs["local"], outs["variance"])


def cEMA(x, alpha):
    """cconsenrich.pyx:5897-5915; n >= 1 and 0 <= alpha <= 1 (outside that the reference is undefined)."""
    is_f32 = isinstance(x, np.ndarray) and x.dtype == np.float32
    arr = np.ascontiguousarray(x, dtype=np.float32 if is_f32 else np.float64).reshape(-1)
    out = np.empty(arr.shape[0], arr.dtype)
    lib = _L()
    if is_f32:
        lib.munc_ema_f32.restype = C.c_int
        lib.munc_ema_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float]
        rc = lib.munc_ema_f32(arr.ctypes.data, out.ctypes.data, arr.shape[0], float(alpha))
    else:
        lib.munc_ema_f64.restype = C.c_int
        lib.munc_ema_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double]
        rc = lib.munc_ema_f64(arr.ctypes.data, out.ctypes.data, arr.shape[0], float(alpha))
    if rc:
        raise ValueError("alpha must lie in [0, 1]")
    return out


# ------------------------------------------------------------------------------------------
# driver-side reductions (oracle/background_oracle.c)
# ------------------------------------------------------------------------------------------
def weighted_mean_residual(stateValues, matrixData, matrixMunc, background=None, pad=0.0):
    """Matrix half of core._relativeSignChangePerKB (core.py:2670-2696): state minus the weighted mean."""
    data = np.ascontiguousarray(matrixData, dtype=np.float32)
    munc = np.ascontiguousarray(matrixMunc, dtype=np.float32)
    state = np.ascontiguousarray(stateValues, dtype=np.float64).reshape(-1)
    bg = None if background is None else np.ascontiguousarray(background, dtype=np.float64).reshape(-1)
    m, n = data.shape
    out = np.empty(n, np.float64)
    lib = _L()
    lib.bg_weighted_mean_residual.restype = None
    lib.bg_weighted_mean_residual.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_double, C.c_void_p]
    lib.bg_weighted_mean_residual(data.ctypes.data, munc.ctypes.data, m, n, state.ctypes.data,
                                  bg.ctypes.data if bg is not None else None, float(pad), out.ctypes.data)
    return out


def sign_change_per_kb(values, intervalSizeBP):
    """core._signChangePerKB (core.py:2614-2644): sign changes per kilobase among the entries that are finite and
    at least 1 % of the mean magnitude."""
    arr = np.asarray(values, dtype=np.float64).reshape(-1)
    if arr.size == 0 or int(intervalSizeBP) <= 0:
        return None
    kept = arr[np.isfinite(arr)]
    if kept.size == 0:
        return None
    floor = 0.01 * float(np.mean(np.abs(kept), dtype=np.float64))
    if not np.isfinite(floor):
        return None
    if floor > 0.0:
        kept = kept[np.abs(kept) >= floor]
    sg = np.sign(kept)
    sg = sg[sg != 0.0]
    flips = int(np.count_nonzero(sg[1:] * sg[:-1] < 0.0)) if sg.size >= 2 else 0
    span_kb = float(arr.size) * float(int(intervalSizeBP)) / 1000.0
    value = float(flips) / span_kb
    return value if np.isfinite(value) else None


def relativeSignChangePerKB(stateValues, matrixData, matrixMunc, *, intervalSizeBP, background=None, pad=0.0):
    """core._relativeSignChangePerKB (core.py:2647-2700) for float32 matrices."""
    return sign_change_per_kb(weighted_mean_residual(stateValues, matrixData, matrixMunc, background, pad), intervalSizeBP)


def interval_diagnostics(covar, munc, obs_prec, q_scale, proc_prec, p_noise, base_q, f, state_dim, cov_init, pad):
    """muncTrace, sumInvR, sumGain0, sumGain1 (float64 [n]) of core._perIntervalOutputDiagnosticTracks
    (core.py:7786-7866); arguments as consenrich_b200.driver.interval_diagnostics."""
    covar = np.ascontiguousarray(covar, dtype=np.float32)
    munc = np.ascontiguousarray(munc, dtype=np.float32)
    n, c = covar.shape[0], covar.shape[1]
    vec = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    obs_prec, q_scale, proc_prec = vec(obs_prec), vec(q_scale), vec(proc_prec)
    p_noise = None if p_noise is None else np.ascontiguousarray(p_noise, dtype=np.float32)
    bq = np.ascontiguousarray(np.asarray(base_q, dtype=np.float64)[:state_dim, :state_dim]).reshape(-1)
    ff = np.ascontiguousarray(np.asarray(f, dtype=np.float64)).reshape(-1) if state_dim == 2 else np.zeros(4)
    outs = [np.empty(n, np.float64) for _ in range(4)]
    lib = _L()
    vp = C.c_void_p
    lib.bg_diag_obs_sums.restype = None
    lib.bg_diag_obs_sums.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_double, vp, vp]
    lib.bg_diag_gain.restype = None
    lib.bg_diag_gain.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp, C.c_double, vp, vp]
    ptr = lambda a: None if a is None else a.ctypes.data
    lib.bg_diag_obs_sums(ptr(munc), munc.shape[0], n, ptr(obs_prec), float(pad), ptr(outs[0]), ptr(outs[1]))
    lib.bg_diag_gain(ptr(covar), ptr(p_noise), ptr(q_scale), ptr(proc_prec), ptr(outs[1]), n, int(state_dim), int(c),
                     ptr(bq), ptr(ff), float(cov_init), ptr(outs[2]), ptr(outs[3]))
    return outs
