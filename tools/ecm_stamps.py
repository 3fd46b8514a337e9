"""Per-tile phase stamps of ONE scan launch inside cb200_ecm_device (diagnostics).

usage: python tools/ecm_stamps.py [launch ...]   (0-based scan launch of the ECM call; even =
forward scans that also compose the smoother's run elements, odd = lean backward scans with the
fused kappa update; default 2 3)"""
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
d, v, _ = bench.synth_device(torch, dev, 1729, m, n, ld)
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True,
                   use_kappa=True)
ts = TrackSweep(m, n, 2, 0, residuals=True)
L = ts.ctx._lib
opts = _lib.EcmOpts(max_iters=bench.ECM_ITERS, inner_iters=bench.T_INNER, update_lambda=0, update_kappa=1,
                    want_outputs=1, init_ones=0, rtol=0.0, nu=bench.ROBUST_NU)
result = _lib.EcmResult()
kap = torch.ones(n, dtype=torch.float32, device=dev)


def ecm():
    kap.fill_(1.0)
    _lib.check(L.cb200_ecm_device(ts.ctx.handle, C.byref(model), C.byref(opts), _p(d), _p(v), m, n, ld, None, None,
                                  _p(kap), _p(ts.xs), _p(ts.Ps), _p(ts.lag), _p(ts.resid), C.byref(result), None))


for _ in range(3):
    ecm()
torch.cuda.synchronize()
T = 6000
for pick in [int(x) for x in (sys.argv[1:] or ["2", "3"])]:
    os.environ["CB200_DEBUG_STAMP_LAUNCH"] = str(pick)
    _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, T, None))
    ecm()
    buf = np.zeros((T, 8), np.int64)
    _lib.check(L.cb200_debug_scan_times(ts.ctx.handle, T, buf.ctypes.data_as(C.c_void_p)))
    k = int((buf[:, 0] > 0).sum())
    b = (buf[:k] - buf[:k, 0].min()) / 1e3
    q = lambda a: np.percentile(a, [0, 50, 100]).round(1).tolist()
    print(f"launch {pick} ({'fwd' if pick % 2 == 0 else 'bwd'}): tiles={k} start{q(b[:,0])} pass1_end{q(b[:,1])} "
          f"prefix{q(b[:,2])} end{q(b[:,3])} | dur pass1{q(b[:,1]-b[:,0])} wait{q(b[:,2]-b[:,1])} "
          f"pass2{q(b[:,3]-b[:,2])}", flush=True)
    for t in list(range(0, k, max(1, k // 16))):
        print(f"  tile {t:4d} start {b[t,0]:6.1f} p1end {b[t,1]:6.1f} aggpub {b[t,7]:6.1f} flagsready {b[t,4]:6.1f} "
              f"loaded {b[t,5]:6.1f} reduced {b[t,6]:6.1f} prefix {b[t,2]:6.1f} end {b[t,3]:6.1f}")
