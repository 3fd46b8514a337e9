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


# ------------------------------------------------------------------------------------------
# cfixedBackgroundECM on a chromosome split into contiguous ranges (lean sweeps, csrc/lean_kernels.cuh)
# ------------------------------------------------------------------------------------------
PAYLOAD = 16  # doubles per shard per pass (include/consenrich_b200.h: CB200_SPLIT_PAYLOAD)


class EcmShard:
    """One shard of a split chromosome on one GPU: the device-side half of ``split_ecm`` (C ABI
    ``cb200_split_*``).  ``data`` / ``munc``: float32 [m, ld] device tensors of the shard's intervals;
    ``kap``: float32 [n] (warm start in, fitted multipliers out); ``qscale``: float32 [n] or None."""

    def __init__(self, ctx, model, nu, data, munc, ld, n, kap, qscale, rank, world, residuals=True):
        import torch
        from . import _lib
        self.torch, self._lib = torch, _lib
        self.ctx, self.model, self.nu = ctx, model, float(nu)
        self.data, self.munc, self.ld, self.n, self.m = data, munc, int(ld), int(n), int(data.shape[0])
        self.kap, self.qs, self.rank, self.world = kap, qscale, int(rank), int(world)
        dev = data.device
        f32, f64 = torch.float32, torch.float64
        self.payload = torch.zeros(PAYLOAD, dtype=f64, device=dev)
        self.sums = torch.zeros(2, dtype=f64, device=dev)
        self.xs = torch.empty((n, 2), dtype=f32, device=dev)
        self.Ps = torch.empty((n, 2, 2), dtype=f32, device=dev)
        self.lag = torch.empty((n, 2, 2), dtype=f32, device=dev)  # row n-1 belongs to a non-final shard only
        self.resid = torch.empty((n, self.m), dtype=f32, device=dev) if residuals else None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _call(self, name, *args):
        self._lib.check(getattr(self.ctx._lib, name)(self.ctx.handle, *args))

    def begin(self):
        self._call("cb200_split_begin", C.byref(self.model), self.nu, self._p(self.data), self._p(self.munc), self.m,
                   self.n, self.ld, self._p(self.qs), self._p(self.kap), int(self.rank == 0),
                   int(self.rank == self.world - 1))
        # the per-pass calls are on the critical path of a sweep at 8 shards (the host issues ~6 calls per
        # sweep): bound once, with their constant arguments already converted
        L, h, vp = self.ctx._lib, self.ctx.handle, C.c_void_p
        self._f = (L.cb200_split_forward_compose, L.cb200_split_forward_replay, L.cb200_split_backward_compose,
                   L.cb200_split_backward_replay)
        self._h = h
        self._pay, self._sums = vp(self.payload.data_ptr()), vp(self.sums.data_ptr())
        self._out = (vp(self.xs.data_ptr()), vp(self.Ps.data_ptr()), vp(self.lag.data_ptr()))
        self._ptr_cache = {}

    def _ptr_of(self, t):
        k = t.data_ptr()
        p = self._ptr_cache.get(k)
        if p is None:
            p = self._ptr_cache[k] = C.c_void_p(k)
        return p

    def forward_compose(self):
        rc = self._f[0](self._h, self._pay)
        if rc:
            self._lib.check(rc)
        return self.payload

    def forward_replay(self, gathered, with_nll, store, track_set):
        rc = self._f[1](self._h, self._ptr_of(gathered), self.rank, self.world, int(with_nll), int(store), int(track_set),
                        self._sums)
        if rc:
            self._lib.check(rc)

    def backward_compose(self, track_set):
        rc = self._f[2](self._h, int(track_set), self._pay)
        if rc:
            self._lib.check(rc)
        return self.payload

    def backward_replay(self, gathered_bwd, gathered_fwd, track_set, publish):
        rc = self._f[3](self._h, self._ptr_of(gathered_bwd), self._ptr_of(gathered_fwd), self.rank, self.world,
                        int(track_set), int(publish), *self._out)
        if rc:
            self._lib.check(rc)

    def end(self):
        if self.resid is not None:
            self._call("cb200_residuals", self._p(self.data), self.m, self.n, self.ld, self._p(self.xs), 2, self._p(self.resid))
        self._call("cb200_split_end", self._p(self.kap))


class LocalGather:
    """The collectives of ``split_ecm`` when every shard lives in this process (one GPU emulating several):
    payloads are stacked, sums added.  ``TorchGather`` is the same interface over NCCL."""

    def gather(self, payloads, out):
        for r, p in enumerate(payloads):
            out[r].copy_(p)
        return out

    def total(self, sums):
        t = sums[0].clone()
        for s in sums[1:]:
            t += s
        return float(t[1].item())


class TorchGather:
    """One shard per rank: ONE all-gather of 128 bytes per rank per pass (NCCL on the GPUs), one all-reduce of
    the NLL per ECM iteration."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group

    def gather(self, payloads, out):
        self.dist.all_gather_into_tensor(out.view(-1), payloads[0], group=self.group)
        return out

    def total(self, sums):
        t = sums[0].clone()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return float(t[1].item())


def split_ecm(shards, comm, max_iters, inner_iters, rtol, patience=2, graphs=None):
    """cfixedBackgroundECM (cconsenrich.pyx:7660-8442; 2-state, kappa the only multiplier fitted) on ONE
    chromosome held as ``shards`` -- the ``EcmShard`` objects of THIS process, in interval order: one per rank
    under torchrun (``comm = TorchGather()``), all of them in a single-process emulation (``LocalGather()``).

    The iteration structure is that of ``cb200_ecm_device`` (csrc/cabi.cu: ecm_device_lean): per iteration
    ``inner_iters`` forward + kappa-carrying backward sweeps, then the NLL forward pass, which -- when another
    iteration follows -- is that iteration's opening pass run ahead into the spare track set; the stopping rule
    is evaluated on the all-reduced NLL, so every rank takes the same decision.  Per pass the shards exchange one
    payload (forward: filtering aggregate + first interval's kappa / qScale; backward: smoothing aggregate + last
    interval's filtered Gaussian).  Returns the diagnostics of the reference's dict (pyx:8409-8425).

    ``graphs``: a dict that lives across calls (or None).  A pass is a fixed sequence of launches -- the shard's
    kernels, the all-gather, the shard's kernels -- with fixed arguments, so with ``graphs`` given every kind of
    pass is captured ONCE into a CUDA graph (stream capture on the shards' stream, NCCL included) and replayed
    afterwards: at 8 shards of chr1 @ 10 bp a sweep is 0.12 ms of kernels per GPU, and ~6 Python-issued calls
    per sweep would otherwise bound it.  The dict must be dropped when the shards (or their sizes) change."""
    torch = shards[0].torch
    world = shards[0].world
    dev = shards[0].data.device
    cache = getattr(shards[0], "_gather_buffers", None)
    if cache is None or len(cache[0]) != len(shards):
        new = lambda: torch.zeros((world, PAYLOAD), dtype=torch.float64, device=dev)
        cache = ([[new(), new()] for _ in shards], [new() for _ in shards], [new() for _ in shards])
        shards[0]._gather_buffers = cache
    # per local shard: the forward payloads each track set was built from; a scratch for forward passes that
    # store nothing; the backward payloads
    g_fwd, g_scr, g_bwd = cache

    def exchange(payloads, outs):
        if len(shards) == 1:
            comm.gather(payloads, outs[0])
        else:  # emulation: every shard sees the same stack
            comm.gather(payloads, outs[0])
            for o in outs[1:]:
                o.copy_(outs[0])

    def forward_eager(track_set, with_nll, store):
        pay = [s.forward_compose() for s in shards]
        outs = [g[track_set] for g in g_fwd] if store else g_scr
        exchange(pay, outs)
        for s, g in zip(shards, outs):
            s.forward_replay(g, with_nll, store, track_set)

    def backward_eager(track_set, publish):
        if not publish:
            pay = [s.backward_compose(track_set) for s in shards]
            exchange(pay, g_bwd)
        for s, gb, gf in zip(shards, g_bwd, g_fwd):
            s.backward_replay(gb, gf[track_set], track_set, publish)

    def replayed(key, fn, *args):
        """Run ``fn(*args)``; with ``graphs`` given, through the CUDA graph of that pass (captured on first use)."""
        if graphs is None:
            return fn(*args)
        g = graphs.get(key)
        if g is None:
            stream = torch.cuda.current_stream(dev)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                fn(*args)
            graphs[key] = g
        g.replay()

    def forward(track_set, with_nll, store):
        replayed(("fwd", track_set, bool(with_nll), bool(store)), forward_eager, track_set, with_nll, store)

    def backward(track_set, publish):
        replayed(("bwd", track_set, bool(publish)), backward_eager, track_set, publish)

    for s in shards:
        s.begin()
    prev, init_nll, has_init, stable, inc, converged = 1.0e16, 0.0, False, 0, 0, False
    rel_impr = abs_rel = 0.0
    cur_set, opened, iters_done = 0, False, 0
    for i in range(int(max_iters)):
        iters_done = i + 1
        for _ in range(int(inner_iters)):
            if not opened:
                forward(cur_set, False, True)
            opened = False
            backward(cur_set, False)
        ahead = i + 1 < max_iters and inner_iters > 0
        forward(1 - cur_set if ahead else cur_set, True, ahead)
        cur = comm.total([s.sums for s in shards])
        # convergence bookkeeping of cconsenrich.pyx:8337-8407
        has_prev = has_init
        if not has_prev:
            init_nll, has_init = cur, True
        elif cur > prev + 1.0e-12 * max(abs(prev), 1.0):
            inc += 1
        delta = abs(cur - prev) if has_prev else 0.0
        scale = max(abs(prev) if has_prev else abs(cur), abs(cur), 1.0)
        rel_impr, abs_rel = ((prev - cur) / scale, delta / scale) if has_prev else (0.0, 0.0)
        prev = cur
        stable = stable + 1 if (has_prev and delta <= rtol * scale) else 0
        if stable >= patience:
            converged = True
            break
        if ahead:
            cur_set, opened = 1 - cur_set, True
    backward(cur_set, True)
    for s in shards:
        s.end()
    return {"iters_done": iters_done, "converged": converged, "stable_iters": stable, "nll_increase_count": inc,
            "initial_nll": init_nll, "final_nll": prev, "final_abs_rel_change": abs_rel,
            "final_rel_improvement": rel_impr}
