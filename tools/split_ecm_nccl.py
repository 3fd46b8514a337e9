"""torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/split_ecm_nccl.py [chrom] [bin_bp] [tracks]

cfixedBackgroundECM on ONE chromosome split over N GPUs (consenrich_b200.sharding.split_ecm: lean sweeps, one
NCCL all-gather of 128 B per rank per pass), checked on every rank against the SAME call done unsharded on that
rank's GPU, and timed against it.  Default: chr1 at 10 bp (24 895 643 intervals, cfg4/cfg5's longest track),
4 tracks.  Prints one JSON line on rank 0."""
import ctypes as C, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from consenrich_b200 import _lib, sharding
from consenrich_b200.device import make_model, _p

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
chrom = sys.argv[1] if len(sys.argv) > 1 else "chr1"
bp = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n = -(-bench.HG38[chrom] // bp)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ld = (n + 31) // 32 * 32
data, munc, _ = bench.synth_device(torch, dev, 1729, m, n, ld)      # the same seeded chromosome on every rank
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True, use_kappa=True)
K, T = bench.ECM_ITERS, bench.T_INNER

# ---- unsharded on this GPU ----
ctx = _lib.Context(local, int(stream.cuda_stream))
L = ctx._lib
f32 = torch.float32
xs, Ps, lag = (torch.empty(s, dtype=f32, device=dev) for s in ((n, 2), (n, 2, 2), (n, 2, 2)))
kap_full = torch.ones(n, dtype=f32, device=dev)
opts = _lib.EcmOpts(max_iters=K, inner_iters=T, update_lambda=0, update_kappa=1, want_outputs=1, init_ones=0, rtol=0.0, nu=bench.ROBUST_NU)
res = _lib.EcmResult()
def whole():
    kap_full.fill_(1.0)
    _lib.check(L.cb200_ecm_device(ctx.handle, C.byref(model), C.byref(opts), _p(data), _p(munc), m, n, ld, None, None,
                                  _p(kap_full), _p(xs), _p(Ps), _p(lag), None, C.byref(res), None))
def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
ms_whole = timed(whole, 3)
nll_whole = float(res.final_nll)

# ---- split: this rank's contiguous range ----
a, b = sharding.split_ranges(n, world)[rank]
nb = b - a
ldb = (nb + 31) // 32 * 32
d_s = torch.zeros((m, ldb), dtype=f32, device=dev); v_s = torch.ones((m, ldb), dtype=f32, device=dev)
d_s[:, :nb] = data[:, a:b]; v_s[:, :nb] = munc[:, a:b]
kap_s = torch.ones(nb, dtype=f32, device=dev)
shard = sharding.EcmShard(_lib.Context(local, int(stream.cuda_stream)), model, bench.ROBUST_NU, d_s, v_s, ldb, nb, kap_s, None,
                          rank, world, residuals=False)
comm = sharding.TorchGather() if world > 1 else sharding.LocalGather()
diag = {}
# SPLIT_GRAPHS=1: replay each pass kind from a CUDA graph.  Measured at 8 GPUs (r2y): 3.96 ms per call against
# 3.94 ms eager -- the pass is bound by the all-gather's latency and the small kernels, not by launches -- so eager
# is the default.  After an NCCL capture torch's process-group teardown does not return; the graph mode therefore
# leaves through os._exit once rank 0 has printed.
graphs = {} if os.environ.get("SPLIT_GRAPHS", "0") == "1" else None
def split():
    kap_s.fill_(1.0)
    diag.update(sharding.split_ecm([shard], comm, max_iters=K, inner_iters=T, rtol=0.0, graphs=graphs))
sharding.split_ecm([shard], comm, max_iters=1, inner_iters=1, rtol=0.0)   # eager once: communicator and arenas exist before any capture
ms_split = timed(split, 5)

# ---- every rank checks its range against the unsharded call ----
def err(g, w):
    g, w = g.double().reshape(len(g), -1), w.double().reshape(len(w), -1)
    return float(((g - w).abs() / w.abs().amax(0, keepdim=True).clamp_min(1e-300)).max())
errs = torch.tensor([err(shard.xs, xs[a:b]), err(shard.Ps, Ps[a:b]), err(kap_s, kap_full[a:b]),
                     err(shard.lag[: nb - (rank == world - 1)], lag[a:b - (rank == world - 1)])], device=dev)
if world > 1: dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    sweeps = K * T
    print(json.dumps({"cuda_graphs": graphs is not None, "what": f"cfixedBackgroundECM (K={K}, t={T}) on {chrom} @ {bp} bp, {m} tracks x {n} intervals, split over {world} GPU(s)",
                      "ms_per_call_unsharded_1gpu": ms_whole, "ms_per_call_split": ms_split,
                      "ms_per_sweep_unsharded": ms_whole / sweeps, "ms_per_sweep_split": ms_split / sweeps,
                      "speedup": ms_whole / ms_split, "parallel_efficiency": ms_whole / ms_split / world,
                      "nll_rel_diff": abs(diag["final_nll"] - nll_whole) / abs(nll_whole),
                      "max_err_over_scale_vs_unsharded": dict(zip(("state", "covariance", "kappa", "lag_covariance"), map(float, errs)))}))
sys.stdout.flush()
if world > 1 and graphs is not None:
    dist.barrier(); torch.cuda.synchronize(); os._exit(0)
if world > 1: dist.destroy_process_group()
