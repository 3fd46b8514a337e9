#!/usr/bin/env python
"""bench.py -- Consenrich state-space hot path on B200: bin.samples filtered+smoothed per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 10-sample ATAC-like count/variance matrices of hg38
chr19 at 25 bp bins (m = 10, n = 2 344 705), 2-state model, per-interval process precision
multipliers on (the CLI default), NLL + residuals + all forward/smoothed tracks emitted.

A *step* = one L-sweep: fold + forward filter + RTS smoother + residuals over one
chromosome, i.e. the work of the reference's cforwardPass + cbackwardPass.  With N ranks every
rank sweeps its own chromosome-sized shard (chromosomes are independent fits: no collective on
the data path), so scaling is weak and `value` is the sum over ranks.

`value`    : device-resident sweeps (inputs already in HBM; 4 rotating replicas = 750 MB > L2).
`e2e`      : the same sweep through the reference-facing host API (consenrich_b200.sweep ->
             cb200_host_sweep): pinned host arrays in, H2D, kernels, D2H of every output track.
`roofline` : dominant kernel's algorithmic bytes / its CUDA-event time inside the timed region.
`cpu_baseline` : the reference's own cforwardPass + cbackwardPass (oracle/_ref, built from the
             unmodified cconsenrich.pyx) on the same matrix, 1 core (its hot path is single-threaded).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_TRACKS = 10
N_BINS = 2_344_705  # ceil(58 617 616 / 25): hg38 chr19 at 25 bp
BIN_BP = 25
METRIC = "bin*samples filtered+smoothed per second (L-sweep: forward filter + RTS smoother + residuals)"
UNIT = "bin*samples/s"
F_MAT = ((1.0, 1.0), (0.0, 1.0))
Q0_MAT = ((1.0e-3, 0.0), (0.0, 1.0e-4))
KAP_BOUNDS = (5.0e-3, 5.0e3)  # constants.py:150-153 (CLI defaults)
N_REPLICAS = 4
WORKLOAD = f"synthetic {M_TRACKS}-sample ATAC, hg38 chr19 @ {BIN_BP} bp ({N_BINS} bins), 2-state, kappa on, residuals on"


# ------------------------------------------------------------------------------------------
# synthetic tracks (SURVEY 8d generator): latent bumps + slow sinusoid, per-sample offset and noise
# ------------------------------------------------------------------------------------------
def synth_host(seed: int, m: int, n: int):
    rng = np.random.default_rng(seed)
    k = np.arange(n, dtype=np.float64)
    x = 0.5 * np.sin(2 * np.pi * k / 5.0e4)
    n_peaks = max(1, n // 800)
    centers = rng.integers(0, n, size=n_peaks)
    widths = rng.uniform(200 / BIN_BP, 2000 / BIN_BP, size=n_peaks)
    heights = rng.uniform(0.5, 4.0, size=n_peaks)
    for c, w, h in zip(centers, widths, heights):
        lo, hi = max(0, int(c - 5 * w)), min(n, int(c + 5 * w))
        x[lo:hi] += h * np.exp(-0.5 * ((k[lo:hi] - c) / w) ** 2)
    v0 = rng.uniform(0.05, 0.3, size=(m, 1))
    munc = (v0 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n))).astype(np.float32)
    data = (x[None, :] + rng.normal(0, 0.05, size=(m, 1)) + rng.standard_normal((m, n)) * np.sqrt(munc)).astype(np.float32)
    kap = np.exp(rng.normal(0.0, 0.5, size=n)).astype(np.float32)
    kap[0] = 1.0
    return np.ascontiguousarray(data), np.ascontiguousarray(munc), kap


def synth_device(torch, dev, seed: int, m: int, n: int, ld: int):
    """Same recipe generated on the device (Philox); rows padded to ld for 16-byte aligned float4 loads."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    k = torch.arange(n, device=dev, dtype=torch.float32)
    x = 0.5 * torch.sin(2 * np.pi * k / 5.0e4)
    n_peaks = max(1, n // 800)
    # bumps via a sparse impulse train smoothed by two box filters (cheap stand-in for Gaussians)
    imp = torch.zeros(n, device=dev)
    idx = torch.randint(0, n, (n_peaks,), device=dev, generator=g)
    imp.index_add_(0, idx, 0.5 + 3.5 * torch.rand(n_peaks, device=dev, generator=g))
    w = 41
    ker = torch.ones(1, 1, w, device=dev) / 8.0
    sm = torch.nn.functional.conv1d(imp.view(1, 1, -1), ker, padding=w // 2)
    sm = torch.nn.functional.conv1d(sm, ker * 8.0 / w, padding=w // 2).view(-1)
    x = x + sm
    v0 = 0.05 + 0.25 * torch.rand(m, 1, device=dev, generator=g)
    data = torch.zeros(m, ld, device=dev)
    munc = torch.ones(m, ld, device=dev)
    munc[:, :n] = v0 * (1.0 + x.abs())[None, :] * (0.5 + torch.rand(m, n, device=dev, generator=g))
    data[:, :n] = (x[None, :] + 0.05 * torch.randn(m, 1, device=dev, generator=g)
                   + torch.randn(m, n, device=dev, generator=g) * munc[:, :n].sqrt())
    kap = torch.exp(0.5 * torch.randn(n, device=dev, generator=g))
    kap[0] = 1.0
    return data, munc, kap


# ------------------------------------------------------------------------------------------
# clocks (sampled during the timed region through NVML)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
def _cpu_module():
    from oracle import oracle as O
    ref = O.load_reference()
    if ref is not None:
        return ref, "reference"
    O.build()
    return O, "port"


def cpu_sweep_seconds(mod, data, munc, kap, reps=1):
    """One L-sweep with the reference's CPU implementation; best of `reps`."""
    m, n = data.shape
    F = np.array(F_MAT, np.float32)
    Q0 = np.array(Q0_MAT, np.float32)
    st = dict(stateForward=np.empty((n, 2), np.float32), stateCovarForward=np.empty((n, 2, 2), np.float32),
              pNoiseForward=np.zeros((n, 2, 2), np.float32), vectorD=np.empty(n, np.float32))
    bw = dict(stateSmoothed=np.empty((n, 2), np.float32), stateCovarSmoothed=np.empty((n, 2, 2), np.float32),
              lagCovSmoothed=np.empty((max(n - 1, 1), 2, 2), np.float32), postFitResiduals=np.empty((n, m), np.float32))
    bm = np.zeros(n, np.int32)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        mod.cforwardPass(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=Q0, intervalToBlockMap=bm,
                         blockCount=1, stateInit=0.0, stateCovarInit=1000.0, pad=1e-4, returnNLL=True,
                         processPrecExp=kap, procPrecisionMultiplierMin=KAP_BOUNDS[0],
                         procPrecisionMultiplierMax=KAP_BOUNDS[1], chunkSize=1000000, **st)
        mod.cbackwardPass(matrixData=data, matrixF=F, stateForward=st["stateForward"],
                          stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"],
                          chunkSize=1000000, **bw)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on the host cores.  The hot path is
    single-threaded per chromosome, so "all the host threads it can use" = independent
    chromosome-sized sweeps, one per thread (the loops release the GIL, cconsenrich.pyx:6578, 6740)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    mod, kind = _cpu_module()
    cores = max(1, min(os.cpu_count() or 1, 64))
    # bounded sample: at most a quarter of chr19 per thread per step, shrunk so that K steps end
    # within ~2 minutes whatever K the driver passes
    probe = synth_host(1, M_TRACKS, 50_000)
    secs_per_bin = cpu_sweep_seconds(mod, *probe, reps=2) / 50_000
    budget = 120.0 / max(args.steps + 1, 1)
    n_sample = int(max(20_000, min(N_BINS // 4, budget / (2.0 * secs_per_bin))))
    data, munc, kap = synth_host(1729, M_TRACKS, n_sample)

    def one(_):
        return cpu_sweep_seconds(mod, data, munc, kap)

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(min(args.warmup, 1)):
            list(ex.map(one, range(cores)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(one, range(cores)))
        dt = time.perf_counter() - t0
    value = M_TRACKS * n_sample * cores * args.steps / dt
    sample = (f"{cores} concurrent sweeps (one per thread) of {M_TRACKS} x {n_sample} bins of the chr19 workload per step, "
              f"cforwardPass+cbackwardPass of {'oracle/_ref (unmodified reference build)' if kind == 'reference' else 'oracle port'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import consenrich_b200 as cb
    from consenrich_b200 import _lib
    from consenrich_b200.device import TrackSweep, make_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # one explicit stream carries everything: the library's kernels, the CUDA events that time them
    # and torch's own work
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    m, n = M_TRACKS, N_BINS
    ld = (n + 31) // 32 * 32
    reps = [synth_device(torch, dev, 1729 + 97 * rank + r, m, n, ld) for r in range(N_REPLICAS)]
    model = make_model(2, F_MAT, Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=KAP_BOUNDS, return_nll=True, use_kappa=True)
    ts = TrackSweep(m, n, 2, local, residuals=True)
    ctx = ts.ctx
    assert ctx.stream_handle == int(stream.cuda_stream) != 0, "library and timing events must share one stream"

    def step(i):
        d, v, kap = reps[i % N_REPLICAS]
        ts.sweep(model, d, v, ld, kap=kap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----
    ctx.reset_timing()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    kern = ctx.kernel_ms()
    ctx.enable_timing(False)
    nll = float(ts.sums[1].item())  # the step's scalar result
    # diagnostic: the same K steps without the per-kernel event pairs (how much the bracketing costs)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms_plain = e0.elapsed_time(e1)

    # ---- e2e through the reference-facing host API, pinned host buffers ----
    host = {}
    d0, v0, k0 = reps[0]
    for key, t in (("data", d0[:, :n]), ("munc", v0[:, :n]), ("kap", k0)):
        h = torch.empty(t.shape, dtype=torch.float32, pin_memory=True)
        h.copy_(t)
        host[key] = h.numpy()
    shapes = dict(stateForward=(n, 2), stateCovarForward=(n, 2, 2), pNoiseForward=(n, 2, 2), vectorD=(n,),
                  stateSmoothed=(n, 2), stateCovarSmoothed=(n, 2, 2), lagCovSmoothed=(n - 1, 2, 2),
                  postFitResiduals=(n, m))
    out = {k_: torch.empty(s, dtype=torch.float32, pin_memory=True).numpy() for k_, s in shapes.items()}
    F = np.array(F_MAT, np.float32)
    Q0 = np.array(Q0_MAT, np.float32)

    def e2e_step():
        return cb.sweep(host["data"], host["munc"], F, Q0, 0.0, 1000.0, pad=1e-4, stateModel=2,
                        processPrecExp=host["kap"], procPrecisionMultiplierMin=KAP_BOUNDS[0],
                        procPrecisionMultiplierMax=KAP_BOUNDS[1], returnNLL=True, wantResiduals=True, out=out)

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    hctx = _lib.default_context(local)
    hl0 = hctx.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_launches = hctx.launch_count - hl0
    h2d = host["data"].nbytes + host["munc"].nbytes + host["kap"].nbytes
    # xf, Pf, D, xs, Ps: n rows; Q and lag-one covariance: n-1 rows; residuals n x m; the two sums
    d2h = n * (8 + 16 + 4 + 8 + 16) + (n - 1) * (16 + 16) + n * m * 4 + 16
    nll_e2e = float(r["sumNLL"])

    # ---- reduce over ranks: max time ----
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    cells = float(m) * float(n)
    value = cells * args.steps * world / (ms_total * 1e-3)
    e2e_value = cells * e2e_steps * world / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback (B200_PROFILING.md)")
        # algorithmic bytes per launch: 8 B per bin*sample read by the fold; 4 + 4 B per
        # bin*sample read + written by the residual pass; per bin the tracks the reference materialises
        # (forward: kappa 4 + xf 8 + Pf 16 + Q 16 + D 4; backward: xf,Pf,Q 40 in, xs,Ps,lag 40 out).
        alg = {"fold": cells * 8.0, "forward_scan": n * 48.0, "backward_scan": n * 80.0,
               "residuals": cells * 8.0 + n * 8.0, "precision_updates": 0.0}
        per = {k_: (v[0] / max(v[1], 1)) for k_, v in kern.items()}  # ms per launch
        dom = max((k_ for k_ in per if kern[k_][1] > 0), key=lambda k_: per[k_])
        achieved = alg[dom] / (per[dom] * 1e-3) / 1e9
        step_ms = ms_total / args.steps
        sweep_alg = cells * 12.0 + n * 84.0  # 12 B per bin*sample + 84 B per bin
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "kernel_ms_per_launch": {k_: per[k_] for k_ in per if kern[k_][1] > 0},
                    "kernel_share_of_step": {k_: per[k_] / step_ms for k_ in per if kern[k_][1] > 0},
                    "sweep": {"algorithmic_bytes": sweep_alg, "achieved": sweep_alg / (step_ms * 1e-3) / 1e9,
                              "frac": sweep_alg / (step_ms * 1e-3) / 1e9 / peak}}
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": None, "sample": None}
        if world == 1 and not args.no_cpu_baseline:
            mod, kind = _cpu_module()
            secs = cpu_sweep_seconds(mod, host["data"], host["munc"], host["kap"], reps=3)
            cpu = {"value": cells / secs, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"the whole workload matrix ({m} x {n}), one L-sweep, best of 3 ({secs:.3f} s)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "m": m, "n": n, "sharding": f"chromosome-sized shard per rank x{world}",
                       "l2": f"inputs larger than L2: {N_REPLICAS} rotating replicas = {N_REPLICAS * 2 * m * ld * 4 / 1e6:.0f} MB"},
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "gpu_launches": int(e2e_launches),
                    "api": "consenrich_b200.sweep -> cb200_host_sweep (pinned host arrays)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "check": {"sum_nll_device": nll, "sum_nll_e2e": nll_e2e,
                      "ms_per_step_without_kernel_events": ms_plain / args.steps},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args)
    else:
        args.steps = 1000 if args.steps is None else args.steps
        args.warmup = 20 if args.warmup is None else max(3, args.warmup)
        run_b200_arm(args)


if __name__ == "__main__":
    main()
