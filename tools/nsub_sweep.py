"""ECM step and plain sweep times against the run length (sub-steps per thread) -- tuning aid.

usage: python tools/nsub_sweep.py [nsub ...]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from consenrich_b200 import _lib
from consenrich_b200.device import TrackSweep, make_model, _p

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
m, n = int(os.environ.get("SWEEP_M", bench.M_TRACKS)), int(os.environ.get("SWEEP_N", bench.N_BINS))
ld = (n + 31) // 32 * 32
reps = [bench.synth_device(torch, dev, 1729 + i, m, n, ld) for i in range(int(os.environ.get("SWEEP_REPS", 4)))]
model = make_model(2, bench.F_MAT, bench.Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=bench.KAP_BOUNDS, return_nll=True,
                   use_kappa=True)
ts = TrackSweep(m, n, 2, 0, residuals=True)
ctx = ts.ctx
L = ctx._lib
opts = _lib.EcmOpts(max_iters=bench.ECM_ITERS, inner_iters=bench.T_INNER, update_lambda=0, update_kappa=1,
                    want_outputs=1, init_ones=0, rtol=0.0, nu=bench.ROBUST_NU)
result = _lib.EcmResult()
kap = torch.ones(n, dtype=torch.float32, device=dev)


def ecm(i=0):
    d, v, _ = reps[i % len(reps)]
    kap.fill_(1.0)
    _lib.check(L.cb200_ecm_device(ctx.handle, C.byref(model), C.byref(opts), _p(d), _p(v), m, n, ld, None, None,
                                  _p(kap), _p(ts.xs), _p(ts.Ps), _p(ts.lag), _p(ts.resid), C.byref(result), None))


def timeit(fn, k=30):
    for i in range(4):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for i in range(k):
        fn(i)
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def fam():
    ctx.reset_timing(); ctx.enable_timing(True)
    for i in range(6):
        ecm(i)
    torch.cuda.synchronize()
    out = {}
    for name, f in (("fold", 0), ("fwd", 1), ("bwd", 2), ("res", 3), ("prec", 4), ("compose", 7), ("segscan", 8), ("publish", 9)):
        ms, cnt = C.c_double(), C.c_int64()
        _lib.check(L.cb200_ctx_kernel_ms(ctx.handle, f, C.byref(ms), C.byref(cnt)))
        if cnt.value:
            out[name] = (round(ms.value / max(cnt.value, 1) * 1e3, 1), cnt.value // 6)
    ctx.enable_timing(False)
    return out


def sweep(i=0):
    d, v, kp = reps[i % len(reps)]
    ts.fold(d, v, ld, model.pad); ts.forward(model, kap=kp); ts.backward(model); ts.residuals(d, ld)


for ns in [int(x) for x in (sys.argv[1:] or ["0"])]:
    _lib.check(L.cb200_set_scan_substeps(ns))
    t = timeit(ecm)
    print(f"nsub={ns:2d} ecm_step_ms={t:.3f} final_nll={result.final_nll:.6f} per-launch us {fam()} "
          f"l_sweep_ms={timeit(sweep):.3f}", flush=True)
