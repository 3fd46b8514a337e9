"""Times pass 1 + warp scan alone (aggregate-only launch) against the full forward / backward scans."""
import ctypes as C, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from consenrich_b200 import _lib
from consenrich_b200.device import TrackSweep, make_model, _p

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
m, n = bench.M_TRACKS, bench.N_BINS
ld = (n + 31) // 32 * 32
d, v, kap = bench.synth_device(torch, dev, 1729, m, n, ld)
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True, use_kappa=True)
ts = TrackSweep(m, n, 2, 0, residuals=True)
L = ts.ctx._lib
agg = torch.zeros(16, dtype=torch.float64, device=dev)
ts.fold(d, v, ld, model.pad)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def fwd_agg():
    _lib.check(L.cb200_forward_shard_aggregate(ts.ctx.handle, C.byref(model), _p(ts.stats), ts.stride, n, None, _p(kap), None, _p(agg)))
def bwd_agg():
    _lib.check(L.cb200_backward_shard_aggregate(ts.ctx.handle, C.byref(model), n, _p(ts.xf), _p(ts.Pf), _p(ts.Qf), 1, _p(agg)))
out = {}
for ns in (4, 8, 11, 16):
    _lib.check(L.cb200_set_scan_substeps(ns))
    ts.forward(model, kap=kap)
    out[ns] = dict(fwd_full=timeit(lambda: ts.forward(model, kap=kap)), fwd_pass1=timeit(fwd_agg),
                   bwd_full=timeit(lambda: ts.backward(model)), bwd_pass1=timeit(bwd_agg),
                   fold=timeit(lambda: ts.fold(d, v, ld, model.pad)), resid=timeit(lambda: ts.residuals(d, ld)))
    print(ns, {k: round(x, 1) for k, x in out[ns].items()}, flush=True)
