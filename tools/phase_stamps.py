"""Per-tile phase stamps of the scan kernels (diagnostics): where a launch spends its time."""
import ctypes as C, os, sys
import numpy as np
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
ts.fold(d, v, ld, model.pad)
for ns in [int(x) for x in (sys.argv[1:] or ["0"])]:
    _lib.check(L.cb200_set_scan_substeps(ns))
    for name, fn in (("fwd", lambda: ts.forward(model, kap=kap)), ("bwd", lambda: ts.backward(model))):
        for _ in range(3): fn()
        T = 6000
        _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, T, None))
        fn()
        buf = np.zeros((T, 8), np.int64)
        _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, T, buf.ctypes.data_as(C.c_void_p)))
        buf = buf[buf[:, 0] > 0]
        t0 = buf[:, 0].min()
        b = (buf - t0) / 1e3
        q = lambda a: np.percentile(a, [0, 50, 100]).round(1).tolist()
        print(f"nsub={ns} {name}: tiles={len(b)} start{q(b[:,0])} pass1_end{q(b[:,1])} prefix{q(b[:,2])} end{q(b[:,3])} | "
              f"dur pass1{q(b[:,1]-b[:,0])} wait{q(b[:,2]-b[:,1])} pass2{q(b[:,3]-b[:,2])}", flush=True)
    if ns == int((sys.argv[1:] or ["0"])[-1]):
        # per-tile series of the last configuration (forward then backward were run; re-run forward)
        _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, 6000, None))
        ts.forward(model, kap=kap)
        buf = np.zeros((6000, 8), np.int64)
        _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, 6000, buf.ctypes.data_as(C.c_void_p)))
        k = int((buf[:, 0] > 0).sum())
        b = (buf[:k] - buf[:k, 0].min()) / 1e3
        for t in list(range(0, 70, 3)) + list(range(70, k, 24)):
            print(f"tile {t:4d} start {b[t,0]:6.1f} p1end {b[t,1]:6.1f} aggpub {b[t,7]:6.1f} flagsready {b[t,4]:6.1f} loaded {b[t,5]:6.1f} reduced {b[t,6]:6.1f} prefix {b[t,2]:6.1f} end {b[t,3]:6.1f}")
