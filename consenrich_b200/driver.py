"""Device versions of the driver-side ``[tracks x intervals]`` reductions of ``consenrich.core``
(SURVEY 8f, next #2).  With the filter / smoother / ECM calls on the GPU these numpy functions are what is
left of ``runConsenrich``'s wall time (profiles/r1p_runconsenrich_e2e.txt).

They are module-level Python functions of ``consenrich.core`` that the driver calls through its module
globals, so they are replaced the way the compiled functions are: by attribute (``install_driver``).  Only
the part of a function that touches the matrices moves to the device; what it then does with per-interval
vectors stays the reference's own code, called through the module it was installed into.

First of them: ``_relativeSignChangePerKB`` (core.py:2647-2700).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .native import _ctx, _ptr

_HOOKS = ("_relativeSignChangePerKB",)
_saved: dict = {}


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
        # argument handling of core.py:2656-2669: anything the device version does not cover goes to the
        # function it replaced
        if stateValues is None or matrixData is None or matrixMunc is None:
            return None
        data, munc = np.asarray(matrixData), np.asarray(matrixMunc)
        if (data.dtype != np.float32 or munc.dtype != np.float32 or data.ndim != 2 or munc.shape != data.shape
                or data.shape[1] != np.asarray(stateValues).size
                or (background is not None and np.asarray(background).size != data.shape[1])):
            return original(stateValues, matrixData, matrixMunc, intervalSizeBP=intervalSizeBP, background=background,
                            pad=pad)
        residual = weighted_mean_residual(stateValues, data, munc, background, pad)
        return module._signChangePerKB(residual, intervalSizeBP=intervalSizeBP)

    return _relativeSignChangePerKB


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
    return module


def uninstall_driver(module=None):
    if module is None:
        import importlib
        module = importlib.import_module("consenrich.core")
    for name, fn in _saved.pop(id(module), {}).items():
        setattr(module, name, fn)
    return module
