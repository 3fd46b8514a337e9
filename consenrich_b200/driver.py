"""Device versions of the driver-side ``[tracks x intervals]`` reductions of ``consenrich.core``
(SURVEY 8f, next #2).  With the filter / smoother / ECM calls on the GPU these numpy functions are what is
left of ``runConsenrich``'s wall time (profiles/r1p_runconsenrich_e2e.txt).

They are module-level Python functions of ``consenrich.core`` that the driver calls through its module
globals, so they are replaced the way the compiled functions are: by attribute (``install_driver``).  Only
the part of a function that touches the matrices moves to the device; what it then does with per-interval
vectors stays the reference's own code, called through the module it was installed into.

Built: ``_relativeSignChangePerKB`` (core.py:2647-2700) and ``_perIntervalOutputDiagnosticTracks``
(core.py:7734-7880, whose per-interval Python loop becomes one thread per interval).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from .native import HostPathWarning, _ctx, _ptr

_HOOKS = ("_relativeSignChangePerKB", "_perIntervalOutputDiagnosticTracks")
_saved: dict = {}


def _host_path(name, why):
    warnings.warn(f"consenrich_b200.driver: {name} ran the reference's host implementation ({why})", HostPathWarning,
                  stacklevel=3)


def weighted_mean_residual(stateValues, matrixData, matrixMunc, background=None, pad=0.0):
    """state - inverse-variance weighted mean over the tracks of (data - background), per interval, float64
    (the matrix half of core.py:2670-2696).  float32 matrices only."""
    data = np.ascontiguousarray(matrixData)
    munc = np.ascontiguousarray(matrixMunc)
    if data.dtype != np.float32 or munc.dtype != np.float32 or data.ndim != 2 or munc.shape != data.shape:
        raise TypeError("matrixData and matrixMunc must be float32 matrices of one shape")
    state = np.ascontiguousarray(stateValues, dtype=np.float64).reshape(-1)
    m, n = data.shape
    if state.shape[0] != n:
        raise ValueError("stateValues length must match the interval count")
    bg = None
    if background is not None:
        bg = np.ascontiguousarray(background, dtype=np.float64).reshape(-1)
        if bg.shape[0] != n:
            raise ValueError("background length must match the interval count")
    out = np.empty(n, np.float64)
    if n:
        ctx = _ctx()
        _lib.check(ctx._lib.cb200_host_weighted_mean_residual(ctx.handle, _ptr(data), _ptr(munc), m, n, _ptr(state),
                                                              _ptr(bg), float(pad), _ptr(out)))
    return out


def _make_relative_sign_change(module, original):
    def _relativeSignChangePerKB(stateValues, matrixData, matrixMunc, *, intervalSizeBP, background=None, pad=0.0):
        # argument handling of core.py:2656-2669.  Malformed shapes go to the function that was replaced, which
        # raises the reference's own errors for them; matrices that are not float32 (runConsenrich always
        # passes float32) are computed by it in their own precision, with a HostPathWarning
        if stateValues is None or matrixData is None or matrixMunc is None:
            return None
        data, munc = np.asarray(matrixData), np.asarray(matrixMunc)
        malformed = (data.ndim != 2 or munc.shape != data.shape or data.shape[1] != np.asarray(stateValues).size
                     or (background is not None and np.asarray(background).size != data.shape[1]))
        if malformed or data.dtype != np.float32 or munc.dtype != np.float32:
            if not malformed:
                _host_path("_relativeSignChangePerKB", f"matrices are {data.dtype}/{munc.dtype}, not float32")
            return original(stateValues, matrixData, matrixMunc, intervalSizeBP=intervalSizeBP, background=background,
                            pad=pad)
        residual = weighted_mean_residual(stateValues, data, munc, background, pad)
        return module._signChangePerKB(residual, intervalSizeBP=intervalSizeBP)

    return _relativeSignChangePerKB


def interval_diagnostics(covar, munc, obs_prec, q_scale, proc_prec, p_noise, base_q, f, state_dim, cov_init, pad):
    """muncTrace, sumInvR, sumGain0, sumGain1 (float64 [n]) of core.py:7786-7866 on the device.
    covar / p_noise: float32 [n, c, c]; munc: float32 [m, n]; the vectors float64 [n]."""
    covar = np.ascontiguousarray(covar, dtype=np.float32)
    munc = np.ascontiguousarray(munc, dtype=np.float32)
    n, cdim = covar.shape[0], covar.shape[1]
    vec = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    obs_prec, q_scale, proc_prec = vec(obs_prec), vec(q_scale), vec(proc_prec)
    if p_noise is not None:
        p_noise = np.ascontiguousarray(p_noise, dtype=np.float32)
        if p_noise.shape != covar.shape:
            raise ValueError("pNoiseForward shape must match stateCovarForward shape")
    outs = [np.empty(n, np.float64) for _ in range(4)]  # muncTrace, sumInvR, sumGain0, sumGain1
    if n == 0:
        return outs
    a = _lib.DiagGainArgs()
    a.covar, a.p_noise = covar.ctypes.data, (p_noise.ctypes.data if p_noise is not None else None)
    a.q_scale, a.proc_prec = q_scale.ctypes.data, (proc_prec.ctypes.data if proc_prec is not None else None)
    a.sum_inv_r, a.sum_gain0, a.sum_gain1 = None, outs[2].ctypes.data, outs[3].ctypes.data
    a.n, a.dim, a.cov_dim, a.cov_init = n, int(state_dim), int(cdim), float(cov_init)
    bq, ff = np.zeros(4), np.zeros(4)
    bq[:state_dim * state_dim] = np.asarray(base_q, dtype=np.float64)[:state_dim, :state_dim].reshape(-1) if state_dim == 1 \
        else np.asarray(base_q, dtype=np.float64)[:2, :2].reshape(-1)
    if state_dim == 2:
        ff[:] = np.asarray(f, dtype=np.float64).reshape(-1)
    a.base_q[:], a.f[:] = list(bq), list(ff)
    ctx = _ctx()
    _lib.check(ctx._lib.cb200_host_interval_diagnostics(ctx.handle, _ptr(munc), munc.shape[0], _ptr(obs_prec), float(pad),
                                                        C.byref(a), _ptr(outs[0]), _ptr(outs[1])))
    return outs


def _make_interval_diagnostics(module, original):
    def _perIntervalOutputDiagnosticTracks(*, stateCovarForward, matrixMunc, matrixQ0, matrixF, stateCovarInit, stateModel,
                                           lambdaExp, processPrecExp, processQScale, pNoiseForward, pad,
                                           obsPrecisionMultiplierMin, obsPrecisionMultiplierMax,
                                           procPrecisionMultiplierMin, procPrecisionMultiplierMax):
        kwargs = dict(stateCovarForward=stateCovarForward, matrixMunc=matrixMunc, matrixQ0=matrixQ0, matrixF=matrixF,
                      stateCovarInit=stateCovarInit, stateModel=stateModel, lambdaExp=lambdaExp,
                      processPrecExp=processPrecExp, processQScale=processQScale, pNoiseForward=pNoiseForward, pad=pad,
                      obsPrecisionMultiplierMin=obsPrecisionMultiplierMin,
                      obsPrecisionMultiplierMax=obsPrecisionMultiplierMax,
                      procPrecisionMultiplierMin=procPrecisionMultiplierMin,
                      procPrecisionMultiplierMax=procPrecisionMultiplierMax)
        covar_in, munc = np.asarray(stateCovarForward), np.asarray(matrixMunc)
        if covar_in.dtype != np.float32 or munc.dtype != np.float32 or (
                pNoiseForward is not None and np.asarray(pNoiseForward).dtype != np.float32):
            # the device version reads float32 tracks, as runConsenrich passes them
            _host_path("_perIntervalOutputDiagnosticTracks", "tracks are not float32")
            return original(**kwargs)
        # shape checks and vector preparation of core.py:7756-7785, 7802-7838 (same texts)
        q0 = np.asarray(matrixQ0, dtype=np.float64)
        f = np.asarray(matrixF, dtype=np.float64)
        mode = module._normalizeStateModel(stateModel)
        dim = 1 if mode == module.STATE_MODEL_LEVEL else 2
        if covar_in.ndim != 3 or covar_in.shape[1] < dim or covar_in.shape[2] < dim:
            raise ValueError("stateCovarForward shape does not match stateModel")
        if covar_in.shape[1] != covar_in.shape[2]:
            _host_path("_perIntervalOutputDiagnosticTracks", "stateCovarForward is not square")
            return original(**kwargs)
        n = int(covar_in.shape[0])
        if munc.ndim != 2 or int(munc.shape[1]) != n:
            raise ValueError("matrixMunc must have shape (trackCount, intervalCount)")
        if q0.ndim != 2 or q0.shape[0] < dim or q0.shape[1] < dim:
            raise ValueError("matrixQ0 shape does not match stateModel")
        if dim == 2 and f.shape != (2, 2):
            raise ValueError("matrixF must have shape (2, 2) for level-trend tracks")
        if lambdaExp is None:
            obs_prec = np.ones(n, dtype=np.float64)
        else:
            obs_prec = np.asarray(lambdaExp, dtype=np.float64).reshape(-1)
            if obs_prec.shape != (n,):
                raise ValueError("lambdaExp length must match interval count")
            obs_prec = np.clip(obs_prec, float(obsPrecisionMultiplierMin), float(obsPrecisionMultiplierMax))
        obs_prec = np.maximum(obs_prec, np.finfo(np.float64).tiny)
        q = module._processQTrackArrays(matrixQ0=q0, intervalCount=n, stateModel=mode, processPrecExp=processPrecExp,
                                        processQScale=processQScale, pNoiseForward=pNoiseForward,
                                        procPrecisionMultiplierMin=float(procPrecisionMultiplierMin),
                                        procPrecisionMultiplierMax=float(procPrecisionMultiplierMax), returnFullQ=False)
        proc_prec = None
        if processPrecExp is not None:
            proc_prec = np.asarray(processPrecExp, dtype=np.float64).reshape(-1)
            proc_prec = np.clip(proc_prec, float(procPrecisionMultiplierMin), float(procPrecisionMultiplierMax))
            proc_prec = np.maximum(proc_prec, np.finfo(np.float64).tiny)
        p_noise = None if pNoiseForward is None or proc_prec is not None else np.asarray(pNoiseForward)
        trace, _, gain0, gain1 = interval_diagnostics(covar_in, munc, obs_prec, q["processQScale"], proc_prec, p_noise,
                                                      q0, f, dim, float(stateCovarInit), float(pad))
        f32 = lambda v: v.astype(np.float32, copy=False)
        return {"baseQLevel": f32(q["baseQLevel"]), "baseQTrend": f32(q["baseQTrend"]),
                "preKappaQLevel": f32(q["preKappaQLevel"]), "preKappaQTrend": f32(q["preKappaQTrend"]),
                "effectiveQLevel": f32(q["effectiveQLevel"]), "effectiveQTrend": f32(q["effectiveQTrend"]),
                "processQScale": f32(q["processQScale"]), "muncTrace": f32(trace), "sumGain0": f32(gain0),
                "sumGain1": f32(gain1)}

    return _perIntervalOutputDiagnosticTracks


def install_driver(module=None):
    """Replace the driver-side reductions of ``consenrich.core`` (or ``module``) that have a device version."""
    if module is None:
        import importlib
        module = importlib.import_module("consenrich.core")
    _lib.load()
    saved = _saved.setdefault(id(module), {})
    if "_relativeSignChangePerKB" not in saved:
        saved["_relativeSignChangePerKB"] = module._relativeSignChangePerKB
    module._relativeSignChangePerKB = _make_relative_sign_change(module, saved["_relativeSignChangePerKB"])
    if "_perIntervalOutputDiagnosticTracks" not in saved:
        saved["_perIntervalOutputDiagnosticTracks"] = module._perIntervalOutputDiagnosticTracks
    module._perIntervalOutputDiagnosticTracks = _make_interval_diagnostics(module,
                                                                           saved["_perIntervalOutputDiagnosticTracks"])
    return module


def uninstall_driver(module=None):
    if module is None:
        import importlib
        module = importlib.import_module("consenrich.core")
    for name, fn in _saved.pop(id(module), {}).items():
        setattr(module, name, fn)
    return module
