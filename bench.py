#!/usr/bin/env python
"""bench.py -- Consenrich state-space hot path on B200: bin.samples filtered+smoothed per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 10-sample ATAC-like count/variance matrices of hg38
chr19 at 25 bp bins (m = 10, n = 2 344 705), 2-state model, per-interval process precision
re-weighting on and observation re-weighting off (the CLI defaults, constants.py:266-282).

A *step* = one L-run of the hot path: ONE call of the reference's native entry point
``cfixedBackgroundECM`` (cconsenrich.pyx:7660) with a fixed iteration budget -- ECM_ITERS x
T_INNER forward-filter + RTS-smoother sweeps with the Student-t kappa update after each, one
NLL-only forward pass per iteration, residuals at the end.  This is what ``runConsenrich`` spends
its time in (it makes several such calls per chromosome).  Throughput counts the smoothed sweeps
only: bin.samples per step = m * n * ECM_ITERS * T_INNER.

With N ranks every rank fits its own chromosome-sized shard (chromosomes are independent fits: no
collective on the data path), so scaling is weak and `value` is the sum over ranks.

`value`    : device-resident steps (tracks already in HBM; 4 rotating replicas = 750 MB > L2).
`e2e`      : the same call through the reference-facing host API (consenrich_b200.cfixedBackgroundECM
             -> cb200_host_ecm): pinned host matrices in, H2D, the whole loop, D2H of every output.
`roofline` : dominant kernel's algorithmic bytes / its CUDA-event time inside the timed region.
`cpu_baseline` : the reference's own cfixedBackgroundECM (oracle/_ref, built from the unmodified
             cconsenrich.pyx) on a bounded sample of the same matrix, 1 core (its hot path is
             single-threaded).
`l_sweep`  : the single forward + backward sweep (fold + filter + smoother + residuals), device-resident
             and through the host API, for reference.
`background` : (N = 1) the rows next to the path at the same size, device-resident and through the host API, with
             the reference's functions timed on one core beside them: background statistics + penalised solve,
             and the variance-stage kernels (seed pass, rolling mean, finalisation).
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
ECM_ITERS = 3       # fixed budget (ECM_fixedBackgroundRtol = 0): the probe of SURVEY 6 saw 3 per call
T_INNER = 5         # constants.py: t_innerIters
SWEEPS_PER_STEP = ECM_ITERS * T_INNER
METRIC = ("bin*samples filtered+smoothed per second (L-run: one cfixedBackgroundECM call = "
          f"{ECM_ITERS} x {T_INNER} filter+smoother sweeps with kappa re-weighting + {ECM_ITERS} NLL passes)")
UNIT = "bin*samples/s"
F_MAT = ((1.0, 1.0), (0.0, 1.0))
Q0_MAT = ((1.0e-3, 0.0), (0.0, 1.0e-4))
KAP_BOUNDS = (5.0e-3, 5.0e3)  # constants.py:150-153 (CLI defaults)
ROBUST_NU = 8.0
N_REPLICAS = 4
WORKLOAD = (f"synthetic {M_TRACKS}-sample ATAC, hg38 chr19 @ {BIN_BP} bp ({N_BINS} bins), 2-state, "
            f"cfixedBackgroundECM iters={ECM_ITERS} (rtol 0) t_inner={T_INNER}, kappa re-weighting on, lambda off")


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


def ecm_kwargs(data, munc):
    """Keyword arguments of one L-run step, identical for the reference and for consenrich_b200."""
    n = data.shape[1]
    return dict(matrixData=data, matrixPluginMuncInit=munc, matrixF=np.array(F_MAT, np.float32),
                matrixQ0=np.array(Q0_MAT, np.float32), intervalToBlockMap=np.zeros(n, np.int32), blockCount=1,
                stateInit=0.0, stateCovarInit=1000.0, ECM_fixedBackgroundIters=ECM_ITERS, ECM_fixedBackgroundRtol=0.0,
                pad=1e-4, ECM_robustTNu=ROBUST_NU, procPrecisionMultiplierMin=KAP_BOUNDS[0],
                procPrecisionMultiplierMax=KAP_BOUNDS[1], ECM_useObsPrecisionReweighting=False,
                ECM_useProcessPrecisionReweighting=True, ECM_useAPN=False, t_innerIters=T_INNER,
                returnIntermediates=True, returnDiagnostics=False, logIterations=False)


def cpu_ecm_seconds(mod, data, munc, reps=1):
    kw = ecm_kwargs(data, munc)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = mod.cfixedBackgroundECM(**kw)
        best = min(best, time.perf_counter() - t0)
    assert out[0] == ECM_ITERS
    return best


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on the host cores.  The hot path is
    single-threaded per chromosome, so "all the host threads it can use" = independent
    chromosome-sized fits, one per thread (the loops release the GIL, cconsenrich.pyx:6578, 6740)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    mod, kind = _cpu_module()
    cores = max(1, min(os.cpu_count() or 1, 64))
    # bounded sample: shrunk so that K steps end within ~2 minutes whatever K the driver passes
    probe = synth_host(1, M_TRACKS, 20_000)
    secs_per_bin = cpu_ecm_seconds(mod, probe[0], probe[1], reps=2) / 20_000
    budget = 120.0 / max(args.steps + min(args.warmup, 1), 1)
    n_sample = int(max(10_000, min(N_BINS // 4, budget / (1.5 * secs_per_bin))))
    data, munc, _ = synth_host(1729, M_TRACKS, n_sample)

    def one(_):
        return cpu_ecm_seconds(mod, data, munc)

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(min(args.warmup, 1)):
            list(ex.map(one, range(cores)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(one, range(cores)))
        dt = time.perf_counter() - t0
    value = M_TRACKS * n_sample * SWEEPS_PER_STEP * cores * args.steps / dt
    sample = (f"{cores} concurrent calls (one per thread) of cfixedBackgroundECM on {M_TRACKS} x {n_sample} bins of the "
              f"chr19 workload per step, {'oracle/_ref (unmodified reference build)' if kind == 'reference' else 'oracle port'}")
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
def background_block(torch, dev, stream, ctx, reps, host, m, n, ld, no_cpu):
    """Times the background-track row: cbackgroundWeightedStatsWithSupport and
    csolveZeroCenteredBackground (zero-centred, lam = 128: constants.py's ECM_backgroundSmoothness) at the
    workload's size, device-resident, through the host API, and the reference's own functions on one core."""
    import ctypes as C
    import consenrich_b200 as cb
    from consenrich_b200 import _lib
    from consenrich_b200.device import _p
    L = ctx._lib
    d, v, _ = reps[0]  # count / variance matrices stand in for residuals / inverse variances
    turn = [0]

    def rotate():  # inputs larger than L2: a different replica every call
        turn[0] += 1
        return reps[turn[0] % len(reps)]
    w = torch.empty(n, dtype=torch.float64, device=dev)
    rhs = torch.empty(n, dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    lam, lam1 = 128.0, 0.0

    def stats():
        dd, vv, _ = rotate()
        _lib.check(L.cb200_background_stats(ctx.handle, _p(dd), _p(vv), m, n, ld, _p(w), _p(rhs), None))

    def solve():
        _lib.check(L.cb200_background_solve(ctx.handle, _p(w), _p(rhs), n, lam, lam1, 1, _p(out), None, None))

    def timed(fn, k=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / k

    ms_stats = timed(stats)
    w.add_(1.0)  # positive weights everywhere: a system the reference accepts
    l0 = ctx.launch_count
    ms_solve = timed(solve, 10)
    launches = (ctx.launch_count - l0) // 13
    hw, hr = w.cpu().numpy(), rhs.cpu().numpy()
    for _ in range(2):
        x = cb.csolveZeroCenteredBackground(hw, hr, lam, True, lamFirst=lam1)
    t0 = time.perf_counter()
    for _ in range(3):
        x = cb.csolveZeroCenteredBackground(hw, hr, lam, True, lamFirst=lam1)
    host_solve = (time.perf_counter() - t0) / 3
    cb.cbackgroundWeightedStatsWithSupport(host["data"], host["munc"])
    t0 = time.perf_counter()
    for _ in range(3):
        cb.cbackgroundWeightedStatsWithSupport(host["data"], host["munc"])
    host_stats = (time.perf_counter() - t0) / 3
    peak = 6458.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    stats_bytes, solve_bytes = 8.0 * m * n + 16.0 * n, 24.0 * n  # in: resid + inv; w, rhs out | in: w, rhs; x out
    blk = {"what": "cbackgroundWeightedStatsWithSupport + csolveZeroCenteredBackground(zeroCenter, lam=128) at the "
                   "workload's size",
           "stats_ms_device": ms_stats, "stats_frac_of_peak": stats_bytes / (ms_stats * 1e-3) / 1e9 / peak,
           "solve_ms_device": ms_solve, "solve_frac_of_peak": solve_bytes / (ms_solve * 1e-3) / 1e9 / peak,
           "solve_kernel_launches": int(launches), "stats_ms_host_api": 1e3 * host_stats,
           "solve_ms_host_api": 1e3 * host_solve, "cpu_reference": None}
    # the rolling local-variance track of the observation-noise stage (cMuncSmoothDenseLocalEvidence), window
    # 41 intervals (~1 kb at 25 bp), per-interval exclusion mask
    window = 41
    mask = (torch.rand(n, device=dev) < 0.02).to(torch.uint8)
    smooth_out = torch.empty((m, ld), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def smooth():
        _, vv, _ = rotate()
        _lib.check(L.cb200_munc_smooth_local_evidence(ctx.handle, _p(vv), _p(mask), 1, m, n, ld, n, window, 1e-12,
                                                      _p(smooth_out), ld, _p(flag)))

    ms_smooth = timed(smooth)
    # cFinalizeMuncEBTrack over one track: local, prior, count floor in, posterior variance out
    fl, fp, fc, fo = (torch.rand(n, device=dev) + 0.01 for _ in range(4))
    fres = _lib.MuncFinalizeResult()

    def finalize():
        _lib.check(L.cb200_munc_finalize_eb(ctx.handle, _p(fl), _p(fp), _p(fc), n, 37.0, 12.0, 1e-3, 5.0, 1, _p(fo),
                                            C.byref(fres)))

    ms_finalize = timed(finalize)  # includes the status read-back (one stream synchronisation per call)
    # cMuncObservationMomentSeedPass, Student-t weights updated, count floor given: 3 matrices in, 4 out
    seed_out = [torch.empty((m, ld), dtype=torch.float32, device=dev) for _ in range(4)]
    seed_cf = torch.rand((m, ld), dtype=torch.float32, device=dev) * 0.1
    seed_vec = [torch.rand(n, dtype=torch.float32, device=dev) * 0.1 for _ in range(2)]
    seed_om = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
    sargs = _lib.MuncSeedArgs()
    sargs.count_floor, sargs.state_mean, sargs.state_var = seed_cf.data_ptr(), seed_vec[0].data_ptr(), seed_vec[1].data_ptr()
    sargs.moment, sargs.rho_out, sargs.local, sargs.variance = (t_.data_ptr() for t_ in seed_out)
    sargs.omega_raw, sargs.omega_out = seed_om[0].data_ptr(), seed_om[1].data_ptr()
    sargs.m, sargs.n, sargs.ld, sargs.active_ld = m, n, ld, n
    sargs.active_mode, sargs.use_weights, sargs.student_t, sargs.update_weights = 0, 1, 1, 1
    sargs.pad, sargs.student_t_df, sargs.d_omega, sargs.omega_min, sargs.omega_max = 1e-4, 8.0, 8.0, 0.01, 100.0
    sargs.variance_floor, sargs.variance_cap = 1e-12, 3.0e38

    def seed():
        dd, vv, _ = rotate()
        sargs.data, sargs.munc = dd.data_ptr(), vv.data_ptr()
        _lib.check(L.cb200_munc_seed_pass(ctx.handle, C.byref(sargs), _p(flag)))

    ms_seed = timed(seed)
    blk["munc_seed_pass_ms_device"] = ms_seed
    blk["munc_seed_pass_frac_of_peak"] = (28.0 * m * n + 16.0 * n) / (ms_seed * 1e-3) / 1e9 / peak
    blk["munc_finalize_ms_device"] = ms_finalize
    blk["munc_finalize_frac_of_peak"] = 16.0 * n / (ms_finalize * 1e-3) / 1e9 / peak
    smooth_bytes = 8.0 * m * n + 1.0 * n
    blk["munc_smooth_ms_device"] = ms_smooth
    blk["munc_smooth_frac_of_peak"] = smooth_bytes / (ms_smooth * 1e-3) / 1e9 / peak
    blk["munc_smooth_window"] = window
    if not no_cpu:
        mod, kind = _cpu_module()
        t0 = time.perf_counter()
        y = mod.csolveZeroCenteredBackground(hw, hr, lam, True, lamFirst=lam1)
        cpu_solve = time.perf_counter() - t0
        t0 = time.perf_counter()
        mod.cbackgroundWeightedStatsWithSupport(host["data"], host["munc"])
        cpu_stats = time.perf_counter() - t0
        hmask = mask.cpu().numpy()
        t0 = time.perf_counter()
        mod.cMuncSmoothDenseLocalEvidence(host["munc"], window, excludeMask=hmask, eps=1e-12)
        cpu_smooth = time.perf_counter() - t0
        t0 = time.perf_counter()
        mod.cMuncObservationMomentSeedPass(host["data"], host["munc"], seed_vec[0].cpu().numpy(), seed_vec[1].cpu().numpy(),
                                           countFloor=np.ascontiguousarray(seed_cf[:, :n].cpu().numpy()))
        cpu_seed = time.perf_counter() - t0
        blk["cpu_reference"] = {"kind": kind, "cores": 1, "solve_ms": 1e3 * cpu_solve, "stats_ms": 1e3 * cpu_stats,
                                "munc_seed_pass_ms": 1e3 * cpu_seed,
                                "munc_smooth_ms": 1e3 * cpu_smooth,
                                "max_abs_diff_over_max_abs": float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))}
    return blk


def run_b200_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import consenrich_b200 as cb
    from consenrich_b200 import _lib
    from consenrich_b200.device import TrackSweep, make_model, _p

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
    L = ctx._lib
    opts = _lib.EcmOpts()
    opts.max_iters, opts.inner_iters, opts.update_lambda, opts.update_kappa, opts.want_outputs = ECM_ITERS, T_INNER, 0, 1, 1
    opts.rtol, opts.nu = 0.0, ROBUST_NU
    result = _lib.EcmResult()
    kap_work = torch.ones(n, dtype=torch.float32, device=dev)

    def step(i):
        d, v, _ = reps[i % N_REPLICAS]
        kap_work.fill_(1.0)  # every step starts from kappa = 1, like a fresh reference call
        _lib.check(L.cb200_ecm_device(ctx.handle, C.byref(model), C.byref(opts), _p(d), _p(v), m, n, ld, None, None,
                                      _p(kap_work), _p(ts.xs), _p(ts.Ps), _p(ts.lag), _p(ts.resid), C.byref(result),
                                      None))

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
    nll = float(result.final_nll)  # the step's scalar result
    assert result.iters_done == ECM_ITERS
    # diagnostic: the same K steps without the per-kernel event pairs (how much the bracketing costs)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms_plain = e0.elapsed_time(e1)

    # ---- secondary: single sweeps (fold + forward + backward + residuals), device-resident ----
    sweep_steps = max(10, min(200, args.steps * 4))
    for i in range(5):
        ts.sweep(model, reps[i % N_REPLICAS][0], reps[i % N_REPLICAS][1], ld, kap=reps[i % N_REPLICAS][2])
    barrier()
    e0.record(stream)
    for i in range(sweep_steps):
        r_ = reps[i % N_REPLICAS]
        ts.sweep(model, r_[0], r_[1], ld, kap=r_[2])
    e1.record(stream)
    barrier()
    ms_sweep = e0.elapsed_time(e1) / sweep_steps

    # ---- e2e through the reference-facing host API, pinned host matrices ----
    host = {}
    d0, v0, k0 = reps[0]
    for key, t in (("data", d0[:, :n]), ("munc", v0[:, :n]), ("kap", k0)):
        h = torch.empty(t.shape, dtype=torch.float32, pin_memory=True)
        h.copy_(t)
        host[key] = h.numpy()
    kw = ecm_kwargs(host["data"], host["munc"])

    def e2e_step():
        return cb.cfixedBackgroundECM(**kw)

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(3):  # keeps two generations of page-locked result arrays alive, as the timed loop does
        r = e2e_step()
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
    h2d = host["data"].nbytes + host["munc"].nbytes                    # tracks (kappa starts at 1 on the device)
    d2h = n * (8 + 16) + (n - 1) * 16 + n * m * 4 + 4 * n + 16          # xs, Ps, lag, residuals, kappa, scalars
    nll_e2e = float(r[1])
    # single sweep through the host API (cforwardPass + cbackwardPass on one upload)
    out = {}
    F = np.array(F_MAT, np.float32)
    Q0 = np.array(Q0_MAT, np.float32)
    for _ in range(2):
        cb.sweep(host["data"], host["munc"], F, Q0, 0.0, 1000.0, processPrecExp=host["kap"],
                 procPrecisionMultiplierMin=KAP_BOUNDS[0], procPrecisionMultiplierMax=KAP_BOUNDS[1], out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        cb.sweep(host["data"], host["munc"], F, Q0, 0.0, 1000.0, processPrecExp=host["kap"],
                 procPrecisionMultiplierMin=KAP_BOUNDS[0], procPrecisionMultiplierMax=KAP_BOUNDS[1], out=out)
    sweep_e2e_s = (time.perf_counter() - t0) / 3

    # ---- background track: statistics + penalised solve, rank 0 at N=1 only ----
    background = None
    if world == 1:
        background = background_block(torch, dev, stream, ctx, reps, host, m, n, ld, args.no_cpu_baseline)

    # ---- reduce over ranks: max time ----
    t = torch.tensor([ms_total, e2e_s, ms_sweep], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, ms_sweep = float(t[0]), float(t[1]), float(t[2])
    cells = float(m) * float(n)
    value = cells * SWEEPS_PER_STEP * args.steps * world / (ms_total * 1e-3)
    e2e_value = cells * SWEEPS_PER_STEP * e2e_steps * world / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback (B200_PROFILING.md)")
        # Algorithmic bytes per launch (SURVEY 8d / DESIGN.md 4): the per-bin tracks each kernel HAS to read
        # and write, averaged over the launches of one step.
        #   forward : fold statistics 32 + kappa 4 in; a storing pass writes xf 8 + Pf 16 + Q 16 (the ECM's
        #             storing passes do not emit D).  Per step: SWEEPS_PER_STEP storing passes (the NLL pass
        #             that closes an iteration is the next iteration's opening pass) + 1 NLL-only pass.
        #   backward: xf, Pf, Q 40 in.  The inner sweeps write only kappa (4): their smoothed tracks feed
        #             nothing but the kappa update, which rides on the replay.  One plain pass per step
        #             writes xs 8 + Ps 16 + lag 16.
        n_fwd, n_bwd = SWEEPS_PER_STEP + 1, SWEEPS_PER_STEP + 1
        fwd_bytes = (SWEEPS_PER_STEP * 76.0 + 36.0) / n_fwd
        bwd_bytes = (SWEEPS_PER_STEP * 44.0 + 80.0) / n_bwd
        alg = {"fold": cells * 8.0 + n * 32.0, "forward_scan": n * fwd_bytes, "backward_scan": n * bwd_bytes,
               "residuals": cells * 8.0 + n * 8.0, "precision_updates": n * 44.0}
        per = {k_: (v[0] / max(v[1], 1)) for k_, v in kern.items()}          # ms per launch
        tot = {k_: v[0] for k_, v in kern.items()}                           # ms inside the timed region
        dom = max((k_ for k_ in tot if kern[k_][1] > 0), key=lambda k_: tot[k_])
        achieved = alg[dom] / (per[dom] * 1e-3) / 1e9
        step_ms = ms_total / args.steps
        # reference-equivalent traffic of the step: the reference re-reads data + munc in every pass
        ref_equiv = SWEEPS_PER_STEP * (cells * 8.0 + n * 84.0) + ECM_ITERS * cells * 8.0 + cells * 4.0
        # DRAM bytes of the same kernel from the committed `ncu --set full` capture (per launch)
        traffic, traffic_src = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1q_ncu_full_summary.json")))
            traffic = float(prof["kernels"][dom]["dram_bytes_per_launch"])
            traffic_src = "profiles/r1q_ncu_full_summary.json (dram__bytes_read.sum + dram__bytes_write.sum)"
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg[dom],
                    "frac_by_kernel": {k_: alg[k_] / (per[k_] * 1e-3) / 1e9 / peak for k_ in per if kern[k_][1] > 0},
                    "kernel_ms_per_launch": {k_: per[k_] for k_ in per if kern[k_][1] > 0},
                    "kernel_launches_per_step": {k_: kern[k_][1] / args.steps for k_ in per if kern[k_][1] > 0},
                    "kernel_share_of_step": {k_: tot[k_] / ms_total for k_ in per if kern[k_][1] > 0},
                    "step": {"reference_equivalent_bytes": ref_equiv,
                             "reference_equivalent_GBps": ref_equiv / (step_ms * 1e-3) / 1e9,
                             "note": "the reference reads data+munc in every pass; here they are folded once per "
                                     "call, so this figure may exceed the HBM peak -- it is not a roofline fraction"}}
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": None, "sample": None}
        if world == 1 and not args.no_cpu_baseline:
            mod, kind = _cpu_module()
            ns = 400_000
            secs = cpu_ecm_seconds(mod, np.ascontiguousarray(host["data"][:, :ns]),
                                   np.ascontiguousarray(host["munc"][:, :ns]), reps=1)
            cpu = {"value": m * ns * SWEEPS_PER_STEP / secs, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"the first {ns} bins of the workload matrix ({m} x {ns}), one cfixedBackgroundECM call "
                             f"({secs:.2f} s)"}
        sweep_alg = cells * 12.0 + n * 84.0  # 12 B per bin*sample + 84 B per bin
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "m": m, "n": n, "sweeps_per_step": SWEEPS_PER_STEP,
                       "sharding": f"chromosome-sized shard per rank x{world}",
                       "l2": f"inputs larger than L2: {N_REPLICAS} rotating replicas = {N_REPLICAS * 2 * m * ld * 4 / 1e6:.0f} MB"},
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "gpu_launches": int(e2e_launches),
                    "api": "consenrich_b200.cfixedBackgroundECM -> cb200_host_ecm (pinned host matrices in, page-locked results out)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "l_sweep": {"what": "one fold + forward filter + RTS smoother + residuals over the same tracks",
                        "ms_per_sweep_device": ms_sweep, "value_device": cells * world / (ms_sweep * 1e-3),
                        "algorithmic_bytes": sweep_alg, "GBps": sweep_alg / (ms_sweep * 1e-3) / 1e9,
                        "frac_of_peak": sweep_alg / (ms_sweep * 1e-3) / 1e9 / peak,
                        "ms_per_sweep_host_api": 1e3 * sweep_e2e_s, "value_host_api": cells / sweep_e2e_s},
            "background": background,
            "check": {"final_nll_device": nll, "final_nll_e2e": nll_e2e,
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
    # stdout carries exactly ONE line, the JSON: libraries that write there on their own (NCCL prints its
    # version line to stdout under torchrun) are sent to stderr for the length of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w", buffering=1)
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args)
    else:
        args.steps = 100 if args.steps is None else args.steps
        args.warmup = 5 if args.warmup is None else max(3, args.warmup)
        run_b200_arm(args)


if __name__ == "__main__":
    main()
