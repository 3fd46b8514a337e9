"""Device-resident sweeps: tracks already in HBM (torch tensors as the carrier), kernels enqueued
on torch's current stream through the C ABI.  This is the path ``bench.py`` times as ``value``.

A sweep = fold + forward filter + RTS smoother (+ residuals), i.e. the work of the reference's
``cforwardPass`` + ``cbackwardPass`` (cconsenrich.pyx:6393, 6635) on one chromosome.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .native import _f32


def _torch():
    import torch
    return torch


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def make_model(state_dim=2, F=((1.0, 1.0), (0.0, 1.0)), Q0=((1e-3, 0.0), (0.0, 1e-4)), state_init=0.0,
               cov_init=1000.0, pad=1e-4, lam_bounds=(0.25, 4.0), kap_bounds=(5e-3, 5e3), return_nll=True,
               store_nll_in_d=False, use_lambda=False, use_kappa=False, use_qscale=False) -> _lib.Model:
    mo = _lib.Model()
    mo.state_dim = int(state_dim)
    F = np.asarray(F, np.float64)
    Q0 = np.asarray(Q0, np.float64)
    mo.F[:] = [F[0, 0], F[0, 1], F[1, 0], F[1, 1]] if state_dim == 2 else [1.0, 0.0, 0.0, 1.0]
    mo.Q0[:] = [Q0[0, 0], Q0[0, 1], Q0[1, 0], Q0[1, 1]] if state_dim == 2 else [Q0[0, 0], 0.0, 0.0, 0.0]
    mo.state_init, mo.cov_init, mo.pad = _f32(state_init), _f32(cov_init), _f32(pad)
    mo.lam_min, mo.lam_max = _f32(lam_bounds[0]), _f32(lam_bounds[1])
    mo.kap_min, mo.kap_max = _f32(kap_bounds[0]), _f32(kap_bounds[1])
    mo.return_nll, mo.store_nll_in_d = int(return_nll), int(store_nll_in_d)
    mo.use_lambda, mo.use_kappa, mo.use_qscale = int(use_lambda), int(use_kappa), int(use_qscale)
    return mo


class TrackSweep:
    """Work tracks for sweeps over one [m x n] chromosome, resident on one GPU."""

    def __init__(self, m: int, n: int, state_dim: int = 2, device: int = 0, residuals: bool = True,
                 ctx: _lib.Context | None = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.CudaError("TrackSweep needs a CUDA device; consenrich_b200 has no CPU path")
        self.m, self.n, self.d = int(m), int(n), int(state_dim)
        self.dev = torch.device("cuda", device)
        # kernels are enqueued on torch's CURRENT stream; the legacy default stream has handle 0, which
        # the C ABI reads as "make your own stream", so it is passed as cudaStreamLegacy (handle 0x1) instead
        self.ctx = ctx or _lib.Context(device, self._stream_handle(torch.cuda.current_stream(self.dev)))
        self.stride = (self.n + 31) // 32 * 32
        f32, f64 = torch.float32, torch.float64
        n, d = self.n, self.d
        self.stats = torch.empty(4 * self.stride, dtype=f64, device=self.dev)
        self.xf = torch.empty((n, d), dtype=f32, device=self.dev)
        self.Pf = torch.empty((n, d, d), dtype=f32, device=self.dev)
        self.Qf = torch.empty((n, d, d), dtype=f32, device=self.dev)
        self.D = torch.empty(n, dtype=f32, device=self.dev)
        self.xs = torch.empty((n, d), dtype=f32, device=self.dev)
        self.Ps = torch.empty((n, d, d), dtype=f32, device=self.dev)
        self.lag = torch.empty((n, d, d), dtype=f32, device=self.dev)  # row n-1 is used by a non-final shard only
        self.resid = torch.empty((n, self.m), dtype=f32, device=self.dev) if residuals else None
        self.sums = torch.zeros(2, dtype=f64, device=self.dev)

    @staticmethod
    def _stream_handle(stream) -> int:
        h = int(stream.cuda_stream)
        return h if h != 0 else 1  # cudaStreamLegacy == (cudaStream_t)0x1

    def bind_current_stream(self):
        self.ctx.set_stream(self._stream_handle(_torch().cuda.current_stream(self.dev)))

    # individual stages (all asynchronous on the context's stream)
    def fold(self, data, munc, ld, pad):
        L = self.ctx._lib
        _lib.check(L.cb200_fold_tracks(self.ctx.handle, _p(data), _p(munc), self.m, self.n, int(ld), float(pad),
                                       _p(self.stats), self.stride))

    def forward(self, model, lam=None, kap=None, qscale=None, store=True, init_state=None):
        L = self.ctx._lib
        _lib.check(L.cb200_forward_scan(
            self.ctx.handle, C.byref(model), _p(self.stats), self.stride, self.m, self.n, _p(lam), _p(kap),
            _p(qscale), _p(init_state), _p(self.xf) if store else None, _p(self.Pf) if store else None,
            _p(self.Qf) if store else None, _p(self.D), _p(self.sums)))

    def backward(self, model, tail_state=None):
        L = self.ctx._lib
        _lib.check(L.cb200_backward_scan(self.ctx.handle, C.byref(model), self.n, _p(self.xf), _p(self.Pf),
                                         _p(self.Qf), _p(tail_state), _p(self.xs), _p(self.Ps), _p(self.lag),
                                         int(self.lag.shape[0])))

    def residuals(self, data, ld):
        L = self.ctx._lib
        _lib.check(L.cb200_residuals(self.ctx.handle, _p(data), self.m, self.n, int(ld), _p(self.xs), self.d,
                                     _p(self.resid)))

    def sweep(self, model, data, munc, ld, lam=None, kap=None, qscale=None):
        """fold + forward + backward (+ residuals): 4 kernel launches (+1 tiny memset per scan)."""
        self.fold(data, munc, ld, model.pad)
        self.forward(model, lam, kap, qscale)
        self.backward(model)
        if self.resid is not None:
            self.residuals(data, ld)
