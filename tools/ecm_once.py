"""One device-resident cb200_ecm_device call at a chosen shape (for ncu captures and launch lists).

usage: SWEEP_M=10 SWEEP_N=2344705 python tools/ecm_once.py [calls]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from consenrich_b200 import _lib
from consenrich_b200.device import TrackSweep, make_model, _p

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
m, n = int(os.environ.get("SWEEP_M", 10)), int(os.environ.get("SWEEP_N", 2344705))
ld = (n + 31) // 32 * 32
d, v, _ = bench.synth_device(torch, dev, 1729, m, n, ld)
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True,
                   use_kappa=True)
ts = TrackSweep(m, n, 2, 0, residuals=True)
ctx = ts.ctx
L = ctx._lib
opts = _lib.EcmOpts(max_iters=bench.ECM_ITERS, inner_iters=bench.T_INNER, update_lambda=0, update_kappa=1,
                    want_outputs=1, init_ones=0, rtol=0.0, nu=bench.ROBUST_NU)
result = _lib.EcmResult()
kap = torch.ones(n, dtype=torch.float32, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    kap.fill_(1.0)
    _lib.check(L.cb200_ecm_device(ctx.handle, C.byref(model), C.byref(opts), _p(d), _p(v), m, n, ld, None, None,
                                  _p(kap), _p(ts.xs), _p(ts.Ps), _p(ts.lag), _p(ts.resid), C.byref(result), None))
torch.cuda.synchronize()
print("final_nll", result.final_nll, "launches", ctx.launch_count)
