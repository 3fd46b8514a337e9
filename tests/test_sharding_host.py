"""Host-side sharding logic, no GPU:

* chromosome assignment (LPT) and contiguous range splitting;
* the SplitSweep protocol over a real 2-process gloo group, with a CPU stand-in for the device
  backend built on tests/emul/scan_emul.cpp (the scan algebra of the kernels compiled by g++),
  checked against the unsharded oracle.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, synth_tracks

HG38_25BP = {  # ceil(len / 25) from src/consenrich/data/hg38.sizes
    "chr1": 9958257, "chr2": 9687742, "chr3": 7931823, "chr4": 7608583, "chr5": 7261531, "chr6": 6832239,
    "chr7": 6373839, "chr8": 5805546, "chr9": 5535789, "chr10": 5351897, "chr11": 5403465, "chr12": 5331012,
    "chr13": 4574574, "chr14": 4281749, "chr15": 4079648, "chr16": 3613534, "chr17": 3330298, "chr18": 3214932,
    "chr19": 2344705, "chr20": 2577767, "chr21": 1868400, "chr22": 2032739, "chrX": 6241636, "chrY": 2289097,
}


def test_lpt_assignment_is_a_balanced_partition():
    from consenrich_b200.sharding import assign_chromosomes
    for world in (1, 2, 4, 8):
        parts = assign_chromosomes(HG38_25BP, world)
        assert len(parts) == world
        flat = [c for p in parts for c in p]
        assert sorted(flat) == sorted(HG38_25BP)  # every chromosome exactly once
        loads = [sum(HG38_25BP[c] for c in p) for p in parts]
        ideal = sum(HG38_25BP.values()) / world
        assert max(loads) <= 1.10 * ideal  # LPT balances the genome to within ~10 % on 8 GPUs
        assert parts == assign_chromosomes(HG38_25BP, world)  # deterministic
    assert assign_chromosomes([5.0, 1.0, 1.0, 1.0, 1.0, 1.0], 2) == [[0], [1, 2, 3, 4, 5]]
    with pytest.raises(ValueError):
        assign_chromosomes(HG38_25BP, 0)


def test_split_ranges_cover_the_chromosome_once():
    from consenrich_b200.sharding import split_ranges
    for n, parts, align in ((9958257, 8, 512), (1000, 3, 512), (5, 5, 512), (4097, 2, 512), (100003, 3, 4)):
        r = split_ranges(n, parts, align)
        assert len(r) == parts and r[0][0] == 0 and r[-1][1] == n
        assert all(a < b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
        if n >= parts * align * 2:
            assert all(a % align == 0 for a, _ in r)
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 2 * align
    with pytest.raises(ValueError):
        split_ranges(3, 4)


# ------------------------------------------------------------------------------------------
# CPU stand-in for DeviceShard: same methods, tensors on the CPU, algebra from scan_emul.cpp
# ------------------------------------------------------------------------------------------
def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmulShard:
    def __init__(self, emul, params, data, munc, lam, kap, qs):
        import torch
        self.torch, self.emul, self.p = torch, emul, params
        self.data, self.munc, self.lam, self.kap, self.qs = data, munc, lam, kap, qs
        self.m, self.n = data.shape
        n = self.n
        self.xf, self.Pf = np.zeros((n, 2), np.float32), np.zeros((n, 2, 2), np.float32)
        self.Qf, self.D = np.zeros((n, 2, 2), np.float32), np.zeros(n, np.float32)
        self.xs, self.Ps = np.zeros((n, 2), np.float32), np.zeros((n, 2, 2), np.float32)
        self.lag = np.zeros((n, 2, 2), np.float32)
        self.Fd = (C.c_double * 4)(1.0, 1.0, 0.0, 1.0)
        self._sums = np.zeros(2)

    def fold(self):
        self.S = [np.empty(self.n, np.float64) for _ in range(4)]
        self.emul.emul_fold(_ptr(self.data), _ptr(self.munc), C.c_int64(self.m), C.c_int64(self.n), C.c_int64(self.n),
                            C.c_double(float(np.float32(1e-4))), *map(_ptr, self.S))

    def forward_aggregate(self):
        agg = np.zeros(16)
        self.emul.emul_forward2_aggregate(_ptr(self.S[0]), _ptr(self.S[1]), C.c_int64(self.n), _ptr(self.lam),
                                          _ptr(self.kap), _ptr(self.qs), C.byref(self.p), _ptr(agg))
        return self.torch.from_numpy(agg)

    def forward_prefix(self, aggs, rank):
        init = np.zeros(8)
        a = np.ascontiguousarray(aggs.numpy())
        self.emul.emul_forward2_prefix(_ptr(a), C.c_int(rank), C.byref(self.p), _ptr(init))
        return init

    def forward_scan(self, init):
        q_head = np.zeros(4, np.float32)
        sd, snll = C.c_double(), C.c_double()
        self.emul.emul_forward2_shard(*map(_ptr, self.S), C.c_int64(self.m), C.c_int64(self.n), _ptr(self.lam),
                                      _ptr(self.kap), _ptr(self.qs), C.byref(self.p), _ptr(self.D), _ptr(self.xf),
                                      _ptr(self.Pf), _ptr(self.Qf), C.byref(sd), C.byref(snll), _ptr(init), _ptr(q_head))
        self._sums[:] = (sd.value, snll.value)
        return self.torch.from_numpy(q_head)

    def set_q_tail(self, q_next):
        self.Qf[self.n - 1] = q_next.numpy().reshape(2, 2)

    def backward_aggregate(self, is_last):
        agg = np.zeros(16)
        self.emul.emul_backward2_aggregate(C.c_int64(self.n), self.Fd, _ptr(self.xf), _ptr(self.Pf), _ptr(self.Qf),
                                           C.c_int(int(is_last)), _ptr(agg))
        return self.torch.from_numpy(agg)

    def backward_prefix(self, saggs, rank, world):
        tail = np.zeros(8)
        a = np.ascontiguousarray(saggs.numpy())
        self.emul.emul_backward2_prefix(_ptr(a), C.c_int(rank), C.c_int(world), _ptr(tail))
        return tail

    def backward_scan(self, tail):
        self.emul.emul_backward2_shard(C.c_int64(self.n), self.Fd, _ptr(self.xf), _ptr(self.Pf), _ptr(self.Qf),
                                       C.byref(self.p), _ptr(self.xs), _ptr(self.Ps), _ptr(self.lag),
                                       C.c_int64(self.n), _ptr(tail))

    def residuals(self):
        self.res = (self.data.T.astype(np.float64) - self.xs[:, :1].astype(np.float64)).astype(np.float32)

    def sums(self):
        return self.torch.from_numpy(self._sums.copy())


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import test_scan_algebra as T
    from consenrich_b200 import sharding
    from oracle import oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emul = C.CDLL(T._EMUL_SO)
        m, n = 5, 30_011
        data, munc = synth_tracks(4242, m, n, masked_frac=0.02)
        rng = np.random.default_rng(7)
        lam = (0.1 + 5 * rng.random(n)).astype(np.float32)
        kap = np.exp(rng.normal(0, 2, n)).astype(np.float32)
        qs = (0.5 + rng.random(n)).astype(np.float32)
        qs[0] = 1.0
        Q0 = np.array([[2e-3, 0.0], [0.0, 1e-4]], np.float32)
        bounds = (0.25, 4.0, 5e-3, 5e3)
        a, b = sharding.split_ranges(n, world, align=512)[rank]
        p = T._params(Q0, 0.25, 8, 32, 99 + rank, lam, kap, qs, bounds, True, False, 1)
        sl = lambda x: np.ascontiguousarray(x[a:b])
        shard = EmulShard(emul, p, np.ascontiguousarray(data[:, a:b]), np.ascontiguousarray(munc[:, a:b]), sl(lam),
                          sl(kap), sl(qs))
        sums = sharding.SplitSweep(shard, sharding.TorchComm()).sweep().numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), a=a, b=b, xf=shard.xf, Pf=shard.Pf, Qf=shard.Qf, D=shard.D,
                 xs=shard.xs, Ps=shard.Ps, lag=shard.lag, res=shard.res, sums=sums)
        if rank == 0:  # unsharded oracle on the whole chromosome
            O.build()
            F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
            w = dict(xf=np.empty((n, 2), np.float32), Pf=np.empty((n, 2, 2), np.float32),
                     Qf=np.zeros((n, 2, 2), np.float32), D=np.empty(n, np.float32))
            r = O.cforwardPass(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=Q0,
                               intervalToBlockMap=np.zeros(n, np.int32), blockCount=1, stateInit=0.25,
                               stateCovarInit=1000.0, pad=1e-4, returnNLL=True, lambdaExp=lam, processPrecExp=kap,
                               processQScale=qs, obsPrecisionMultiplierMin=bounds[0],
                               obsPrecisionMultiplierMax=bounds[1], procPrecisionMultiplierMin=bounds[2],
                               procPrecisionMultiplierMax=bounds[3], stateForward=w["xf"], stateCovarForward=w["Pf"],
                               pNoiseForward=w["Qf"], vectorD=w["D"])
            bw = O.cbackwardPass(matrixData=data, matrixF=F, stateForward=w["xf"], stateCovarForward=w["Pf"],
                                 pNoiseForward=w["Qf"])
            np.savez(os.path.join(out_dir, "oracle.npz"), nll=r[3], phi=r[0], xs=bw[0], Ps=bw[1], lag=bw[2], res=bw[3], **w)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_split_sweep_over_gloo_matches_the_unsharded_oracle(tmp_path, oracle):
    import socket

    import torch.multiprocessing as mp

    import test_scan_algebra as T
    from parity_util import assert_sweep_tracks_close as close

    # build the emulation library once, in the parent
    hdr = os.path.join(ROOT, "consenrich_b200", "csrc", "ssm_math.cuh")
    if (not os.path.exists(T._EMUL_SO)) or os.path.getmtime(T._EMUL_SO) < max(os.path.getmtime(T._EMUL_SRC),
                                                                               os.path.getmtime(hdr)):
        import subprocess
        os.makedirs(os.path.dirname(T._EMUL_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off",
                               T._EMUL_SRC, "-o", T._EMUL_SO])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = np.load(tmp_path / "oracle.npz")
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert parts[0]["a"] == 0 and parts[0]["b"] == parts[1]["a"]
    n = int(parts[-1]["b"])
    cat = lambda k: np.concatenate([p[k] for p in parts])
    close(cat("xf"), want["xf"], "stateForward")
    close(cat("Pf"), want["Pf"], "stateCovarForward", scale="component")
    np.testing.assert_array_equal(cat("Qf")[: n - 1], want["Qf"][: n - 1])  # includes the exchanged boundary row
    close(cat("D"), want["D"], "vectorD")
    close(cat("xs"), want["xs"], "stateSmoothed")
    close(cat("Ps"), want["Ps"], "stateCovarSmoothed", scale="component")
    close(cat("lag")[: n - 1], want["lag"], "lagCovSmoothed", scale="component")
    close(cat("res"), want["res"], "postFitResiduals")
    for p in parts:  # the all-reduce handed every rank the chromosome's totals
        assert abs(p["sums"][1] - want["nll"]) <= 2e-6 * abs(want["nll"])
        assert abs(np.float32(p["sums"][0] / n) - want["phi"]) <= 1e-4 * abs(want["phi"])


# ------------------------------------------------------------------------------------------
# split_ecm over gloo: the exchange protocol and the (replicated) stopping rule, with a recording stand-in
# for the device shard
# ------------------------------------------------------------------------------------------
class RecordingShard:
    """Stands in for sharding.EcmShard: payloads carry (rank, pass counter), the NLL part of each shard is a
    fixed sequence that converges, and every call is logged with what it was handed."""

    def __init__(self, rank, world):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        self.data = torch.zeros(1)  # split_ecm reads .device from it
        self.sums = torch.zeros(2, dtype=torch.float64)
        self.payload = torch.zeros(16, dtype=torch.float64)
        self.log, self.passes, self.nll_calls = [], 0, 0

    def begin(self):
        self.log.append(("begin",))

    def forward_compose(self):
        self.passes += 1
        self.payload[:] = 0
        self.payload[0], self.payload[1], self.payload[14] = self.rank, self.passes, 100 + self.rank
        return self.payload

    def forward_replay(self, gathered, with_nll, store, track_set):
        self.log.append(("fwd", int(track_set), bool(with_nll), bool(store), gathered.clone()))
        if with_nll:
            self.nll_calls += 1
            self.sums[1] = (1000.0 + 100.0 * 0.1 ** self.nll_calls) * (self.rank + 1)  # converges geometrically

    def backward_compose(self, track_set):
        self.passes += 1
        self.payload[:] = 0
        self.payload[0], self.payload[1], self.payload[9] = self.rank, self.passes, 200 + self.rank
        return self.payload

    def backward_replay(self, gathered_bwd, gathered_fwd, track_set, publish):
        self.log.append(("bwd", int(track_set), bool(publish), gathered_bwd.clone(), gathered_fwd.clone()))

    def end(self):
        self.log.append(("end",))


def _ecm_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import pickle

    import torch.distributed as dist

    from consenrich_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = RecordingShard(rank, world)
        diag = sharding.split_ecm([shard], sharding.TorchGather(), max_iters=12, inner_iters=2, rtol=1e-4)
        log = [tuple(x.numpy() if hasattr(x, "numpy") else x for x in e) for e in shard.log]
        with open(os.path.join(out_dir, f"ecm{rank}.pkl"), "wb") as f:
            pickle.dump((diag, log), f)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_split_ecm_protocol_over_gloo(tmp_path):
    import pickle
    import socket

    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_ecm_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    runs = [pickle.load(open(tmp_path / f"ecm{r}.pkl", "rb")) for r in range(world)]
    (d0, l0), (d1, l1) = runs
    # every rank saw the same all-reduced NLL, hence took the same decisions
    assert d0 == d1 and d0["converged"] and 2 < d0["iters_done"] < 12
    assert [e[:3] if e[0] == "fwd" else e[:3] for e in l0] == [e[:3] if e[0] == "fwd" else e[:3] for e in l1]
    assert l0[0] == ("begin",) and l0[-1] == ("end",) and l0[-2][0] == "bwd" and l0[-2][2] is True  # publish last
    for log in (l0, l1):
        stored = {}  # track set -> forward payloads it was built from
        for e in log[1:-1]:
            if e[0] == "fwd":
                _, s, with_nll, store, g = e
                assert g.shape == (world, 16) and list(g[:, 0]) == [0.0, 1.0]      # rank order
                assert g[0, 1] == g[1, 1] and list(g[:, 14]) == [100.0, 101.0]     # same pass on both ranks
                if store:
                    stored[s] = g
            else:
                _, s, publish, gb, gf = e
                # the backward pass of a track set is handed the forward payloads of the pass that stored it,
                # even when a later, non-storing NLL pass has run in between
                assert np.array_equal(gf, stored[s])
                if not publish:
                    assert list(gb[:, 0]) == [0.0, 1.0] and list(gb[:, 9]) == [200.0, 201.0]
    # iteration structure: 2 sweeps per iteration, the closing NLL pass opens the next iteration
    kinds = [(e[0],) + tuple(e[1:4] if e[0] == "fwd" else e[1:3]) for e in l0[1:-1]]
    assert kinds[:5] == [("fwd", 0, False, True), ("bwd", 0, False), ("fwd", 0, False, True), ("bwd", 0, False),
                         ("fwd", 1, True, True)]
    assert kinds[5:8] == [("bwd", 1, False), ("fwd", 1, False, True), ("bwd", 1, False)]
