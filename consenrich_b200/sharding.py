"""Sharding of the state-space hot path across GPUs.

Two levels, one process per GPU:

* **chromosomes** are independent fits (separate ``runConsenrich`` calls in the reference,
  consenrich.py:8809): ``assign_chromosomes`` bin-packs them onto ranks by cost (longest
  processing time first).  No communication at all.
* **a chromosome too long for its share** is split into contiguous bin ranges, one per rank
  (``split_ranges``).  ``SplitSweep`` then runs one forward + backward sweep with exactly three tiny
  exchanges: the per-shard filtering aggregates (14 float64 each), the process-noise row that
  straddles each boundary (4 float32), and the per-shard smoothing aggregates (9 float64); plus a
  sum of the two scalars (sum D, sum NLL).  Everything else is shard-local.

The collective is whatever ``torch.distributed`` group is passed (NCCL on the GPUs; the CPU tests
run the same protocol over gloo with a CPU stand-in for the device backend).  The payloads are a
few hundred bytes, so the exchange is latency-bound and a plain all-gather is the right tool.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

AGG_PITCH = 16   # doubles per gathered aggregate (14 / 9 / 5 / 3 used)
STATE_PITCH = 8  # doubles per prefix state (5 / 2 used)


# ------------------------------------------------------------------------------------------
# chromosome-level sharding: no communication
# ------------------------------------------------------------------------------------------
def assign_chromosomes(costs: Mapping[str, float] | Sequence[float], world: int) -> list[list]:
    """Longest-processing-time-first bin packing.  ``costs``: {name: cost} or a sequence of costs
    (cost ~ tracks x bins).  Returns, per rank, the names (or indices) assigned to it, in decreasing
    cost order.  Deterministic: ties break on the name / index."""
    if world <= 0:
        raise ValueError("world must be positive")
    items = list(costs.items()) if isinstance(costs, Mapping) else list(enumerate(costs))
    items.sort(key=lambda kv: (-float(kv[1]), str(kv[0])))
    load = [0.0] * world
    out: list[list] = [[] for _ in range(world)]
    for name, cost in items:
        r = min(range(world), key=lambda i: (load[i], i))
        out[r].append(name)
        load[r] += float(cost)
    return out


def split_ranges(n: int, parts: int, align: int = 512) -> list[tuple[int, int]]:
    """Contiguous [start, stop) bin ranges covering [0, n), boundaries on multiples of ``align``
    (the scan kernels' tile granule) where possible; every range is non-empty when n >= parts."""
    if parts <= 0:
        raise ValueError("parts must be positive")
    if n < parts:
        raise ValueError("cannot split fewer intervals than shards")
    bounds = [0]
    for p in range(1, parts):
        b = (n * p // parts) // align * align
        b = max(b, bounds[-1] + 1)
        b = min(b, n - (parts - p))
        bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]


# ------------------------------------------------------------------------------------------
# communication shims
# ------------------------------------------------------------------------------------------
class TorchComm:
    """all_gather / all_reduce of small tensors over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, t):
        import torch
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t


# ------------------------------------------------------------------------------------------
# one sweep of a chromosome split into contiguous ranges
# ------------------------------------------------------------------------------------------
class SplitSweep:
    """Forward filter + RTS smoother of ONE chromosome whose bins are split across the ranks of
    ``comm``; this rank holds bins [start, stop).  ``backend`` does the shard-local device work
    (``DeviceShard`` below on a GPU).  Protocol, identical on every rank:

        1. fold the local tracks, reduce the range to one filtering element        (local)
        2. all-gather the elements, apply the ordered prefix of ranks < r          (128 B / rank)
        3. forward scan from that state; all-gather the boundary Q rows            (16 B / rank)
        4. reduce the local forward tracks to one smoothing element, all-gather,
           apply the ordered suffix of ranks > r                                   (72 B / rank)
        5. backward scan from that tail state, residuals, sum of the two scalars   (16 B)
    """

    def __init__(self, backend, comm):
        self.b = backend
        self.comm = comm

    def sweep(self):
        b, comm = self.b, self.comm
        r, w = comm.rank, comm.world
        b.fold()
        aggs = comm.all_gather(b.forward_aggregate())               # [w, AGG_PITCH] float64
        init = b.forward_prefix(aggs, r) if r > 0 else None
        q_head = b.forward_scan(init)                               # float32 [d*d]: Q of this shard's first bin
        heads = comm.all_gather(q_head)                             # [w, d*d]
        if r + 1 < w:
            b.set_q_tail(heads[r + 1])                              # row n-1 of this shard's pNoiseForward
        saggs = comm.all_gather(b.backward_aggregate(is_last=(r == w - 1)))
        tail = b.backward_prefix(saggs, r, w) if r + 1 < w else None
        b.backward_scan(tail)
        b.residuals()
        return comm.all_reduce_sum(b.sums())                        # [sum D, sum NLL] over the chromosome


def run_split_local(backends):
    """The SplitSweep protocol for all shards inside ONE process, phase by phase (the shards'
    backends in bin order).  What ``torchrun`` does with one shard per rank; used to exercise the
    shard entry points of the C ABI on a single GPU.  Returns [sum D, sum NLL] of the chromosome."""
    import torch
    w = len(backends)
    for b in backends:
        b.fold()
    aggs = torch.stack([b.forward_aggregate().clone() for b in backends])
    heads = torch.stack([b.forward_scan(b.forward_prefix(aggs, r) if r > 0 else None).clone()
                         for r, b in enumerate(backends)])
    for r, b in enumerate(backends[:-1]):
        b.set_q_tail(heads[r + 1])
    saggs = torch.stack([b.backward_aggregate(is_last=(r == w - 1)).clone() for r, b in enumerate(backends)])
    total = None
    for r, b in enumerate(backends):
        b.backward_scan(b.backward_prefix(saggs, r, w) if r + 1 < w else None)
        b.residuals()
        total = b.sums() if total is None else total + b.sums()
    return total


class DeviceShard:
    """Shard-local device work for SplitSweep through the C ABI (tracks already in HBM)."""

    def __init__(self, ts, model, data, munc, ld, lam=None, kap=None, qscale=None):
        import torch
        self.torch = torch
        self.ts, self.model = ts, model
        self.data, self.munc, self.ld = data, munc, int(ld)
        self.lam, self.kap, self.qs = lam, kap, qscale
        dev = ts.dev
        d = ts.d
        self.agg = torch.zeros(AGG_PITCH, dtype=torch.float64, device=dev)
        self.sagg = torch.zeros(AGG_PITCH, dtype=torch.float64, device=dev)
        self.init = torch.zeros(STATE_PITCH, dtype=torch.float64, device=dev)
        self.tail = torch.zeros(STATE_PITCH, dtype=torch.float64, device=dev)
        self.q_head = torch.zeros(d * d, dtype=torch.float32, device=dev)

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _call(self, name, *args):
        from . import _lib
        _lib.check(getattr(self.ts.ctx._lib, name)(self.ts.ctx.handle, *args))

    def fold(self):
        self.ts.fold(self.data, self.munc, self.ld, self.model.pad)

    def forward_aggregate(self):
        ts = self.ts
        self._call("cb200_forward_shard_aggregate", C.byref(self.model), self._p(ts.stats), ts.stride, ts.n,
                   self._p(self.lam), self._p(self.kap), self._p(self.qs), self._p(self.agg))
        return self.agg

    def forward_prefix(self, aggs, rank):
        self._call("cb200_forward_shard_prefix", C.byref(self.model), self._p(aggs.contiguous()), int(rank),
                   self._p(self.init))
        return self.init

    def forward_scan(self, init):
        ts = self.ts
        self._call("cb200_forward_scan_shard", C.byref(self.model), self._p(ts.stats), ts.stride, ts.m, ts.n,
                   self._p(self.lam), self._p(self.kap), self._p(self.qs), self._p(init), self._p(ts.xf),
                   self._p(ts.Pf), self._p(ts.Qf), self._p(ts.D), self._p(ts.sums), self._p(self.q_head))
        return self.q_head

    def set_q_tail(self, q_next):
        self.ts.Qf[self.ts.n - 1].view(-1).copy_(q_next.view(-1))

    def backward_aggregate(self, is_last):
        ts = self.ts
        self._is_last = bool(is_last)
        self._call("cb200_backward_shard_aggregate", C.byref(self.model), ts.n, self._p(ts.xf), self._p(ts.Pf),
                   self._p(ts.Qf), int(bool(is_last)), self._p(self.sagg))
        return self.sagg

    def backward_prefix(self, saggs, rank, world):
        self._call("cb200_backward_shard_prefix", C.byref(self.model), self._p(saggs.contiguous()), int(rank),
                   int(world), self._p(self.tail))
        return self.tail

    def backward_scan(self, tail):
        self.ts.backward(self.model, tail_state=tail)

    def residuals(self):
        if self.ts.resid is not None:
            self.ts.residuals(self.data, self.ld)

    def sums(self):
        return self.ts.sums.clone()
