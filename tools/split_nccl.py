"""torchrun --nproc-per-node N tools/split_nccl.py : one chromosome split over N GPUs (NCCL), checked
against the same sweep done unsharded on rank 0's GPU."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from consenrich_b200 import sharding
from consenrich_b200.device import TrackSweep, make_model

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
m, n = 10, 2_344_705
data, munc, kap = bench.synth_host(1729, m, n)          # same seeded chromosome on every rank
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True, use_kappa=True)
a, b = sharding.split_ranges(n, world)[rank]
def dev_tracks(x, lo, hi):
    nb = hi - lo; ld = (nb + 31) // 32 * 32
    t = torch.ones((m, ld), dtype=torch.float32, device=dev); t[:, :nb] = torch.from_numpy(x[:, lo:hi]).to(dev)
    return t, ld
d, ld = dev_tracks(data, a, b); v, _ = dev_tracks(munc, a, b)
ts = TrackSweep(m, b - a, 2, local, residuals=True)
shard = sharding.DeviceShard(ts, model, d, v, ld, kap=torch.from_numpy(kap[a:b].copy()).to(dev))
split = sharding.SplitSweep(shard, sharding.TorchComm())
for _ in range(3): sums = split.sweep()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20): sums = split.sweep()
e1.record(stream); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# gather the smoothed level on rank 0 and compare with the unsharded sweep
xs_parts = [torch.empty((bb - aa, 2), dtype=torch.float32, device=dev) for aa, bb in sharding.split_ranges(n, world)]
Ps_parts = [torch.empty((bb - aa, 2, 2), dtype=torch.float32, device=dev) for aa, bb in sharding.split_ranges(n, world)]
if world > 1:
    for r, (aa, bb) in enumerate(sharding.split_ranges(n, world)):
        src_x = ts.xs if r == rank else xs_parts[r]; src_p = ts.Ps if r == rank else Ps_parts[r]
        dist.broadcast(src_x, src=r); dist.broadcast(src_p, src=r)
        xs_parts[r], Ps_parts[r] = src_x, src_p
else:
    xs_parts, Ps_parts = [ts.xs], [ts.Ps]
if rank == 0:
    dfull, ldf = dev_tracks(data, 0, n); vfull, _ = dev_tracks(munc, 0, n)
    tf = TrackSweep(m, n, 2, local, residuals=False)
    tf.sweep(model, dfull, vfull, ldf, kap=torch.from_numpy(kap).to(dev)); torch.cuda.synchronize()
    xs = torch.cat(xs_parts).cpu().numpy().astype(np.float64); want = tf.xs.cpu().numpy().astype(np.float64)
    Ps = torch.cat(Ps_parts).cpu().numpy().astype(np.float64); wantP = tf.Ps.cpu().numpy().astype(np.float64)
    ex = np.abs(xs - want).max() / np.abs(want).max()
    eP = (np.abs(Ps - wantP).reshape(n, 4).max(0) / np.abs(wantP).reshape(n, 4).max(0)).max()
    nll_full = float(tf.sums[1]); nll_split = float(sums[1])
    print(f"world={world} split sweep {float(ms):.3f} ms/sweep; max rel err vs unsharded: state {ex:.2e}, cov {eP:.2e}; "
          f"NLL split {nll_split:.6f} vs unsharded {nll_full:.6f} (rel {abs(nll_split-nll_full)/abs(nll_full):.1e})", flush=True)
    assert ex < 1e-5 and eP < 1e-4 and abs(nll_split - nll_full) <= 2e-6 * abs(nll_full)
dist.destroy_process_group()
