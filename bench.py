#!/usr/bin/env python
"""bench.py -- Consenrich state-space hot path on B200: bin.samples filtered+smoothed per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|cfg4|cfg5] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; synthetic count / variance matrices, hg38 chromosome lengths):
    cfg2  10 samples, chr19 at 25 bp (2 344 705 bins)                       -- one chromosome
    cfg3  50 samples, the 24 hg38 chromosomes at 25 bp (123 530 804 bins)    -- DEFAULT: the configuration the
          metric is quoted on ("1/2/4/8 B200"); the whole genome fits one B200 (49 GB of tracks)
    cfg4  200 samples, whole genome at 10 bp (308 826 993 bins)
    cfg5  1000 samples, whole genome at 50 bp (61 765 409 bins)
2-state model, per-interval process precision re-weighting on and observation re-weighting off (the CLI
defaults, constants.py:266-282).

A *step* = one pass of the hot path over the WHOLE genome of the configuration: for every chromosome ONE
call of the reference's native entry point ``cfixedBackgroundECM`` (cconsenrich.pyx:7660) with a fixed
iteration budget -- ECM_ITERS x T_INNER forward-filter + RTS-smoother sweeps with the Student-t kappa
update after each, one NLL-only forward pass per iteration, smoothed tracks + residuals at the end.  This
is what ``runConsenrich`` spends its time in.  Throughput counts the smoothed sweeps only:
bin.samples per step = m * (genome bins) * ECM_ITERS * T_INNER.

With N ranks the chromosomes are bin-packed onto the ranks (longest first; chromosomes are independent
fits, so there is NO collective on the data path) and every rank works through its own list: the total
work is fixed, i.e. STRONG scaling; the time of a step is the slowest rank's.

`value`    : device-resident steps (tracks already in HBM).
`e2e`      : the same genome pass through the reference-facing host API (consenrich_b200.cfixedBackgroundECM
             -> cb200_host_ecm): pinned host matrices in, H2D, the whole loop, D2H of every output.
`roofline` : the dominant kernel's algorithmic bytes / its CUDA-event time inside the timed region.
`cpu_baseline` : the reference's own cfixedBackgroundECM (oracle/_ref, the unmodified cconsenrich.pyx built
             in-tree) on a bounded slice of the same matrix, 1 core (its hot path is single-threaded).
`check`    : the GPU call on that same slice compared with the reference's outputs.
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

# hg38 primary chromosomes (UCSC hg38.chrom.sizes; the reference ships the same table as
# src/consenrich/data/hg38.sizes)
HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}
CONFIGS = {
    "cfg2": dict(m=10, bin_bp=25, chroms=["chr19"], assay="ATAC"),
    "cfg3": dict(m=50, bin_bp=25, chroms=list(HG38), assay="ChIP-seq"),
    "cfg4": dict(m=200, bin_bp=10, chroms=list(HG38), assay="DNase-seq"),
    "cfg5": dict(m=1000, bin_bp=50, chroms=list(HG38), assay="CUT&RUN"),
}
ECM_ITERS = 3       # fixed budget (ECM_fixedBackgroundRtol = 0): the probe of SURVEY 6 saw 3 per call
T_INNER = 5         # constants.py: t_innerIters
SWEEPS_PER_STEP = ECM_ITERS * T_INNER
METRIC = ("bin*samples filtered+smoothed per second (L-run over the configuration's genome: per chromosome one "
          f"cfixedBackgroundECM call = {ECM_ITERS} x {T_INNER} filter+smoother sweeps with kappa re-weighting + "
          f"{ECM_ITERS} NLL passes)")
UNIT = "bin*samples/s"
F_MAT = ((1.0, 1.0), (0.0, 1.0))
Q0_MAT = ((1.0e-3, 0.0), (0.0, 1.0e-4))
KAP_BOUNDS = (5.0e-3, 5.0e3)  # constants.py:150-153 (CLI defaults)
ROBUST_NU = 8.0
# kept for the probes under tools/ (cfg2's single chromosome)
M_TRACKS, N_BINS, BIN_BP = 10, 2_344_705, 25
RESIDENT_BUDGET = 120e9  # bytes of tracks kept resident per GPU; beyond it chromosomes share rotating buffers


def chrom_bins(cfg):
    bp = cfg["bin_bp"]
    return {c: -(-HG38[c] // bp) for c in cfg["chroms"]}


def make_config(name):
    """The `config` object of the JSON line: identical for both arms (what is measured, not how)."""
    cfg = CONFIGS[name]
    bins = chrom_bins(cfg)
    return {
        "workload": (f"{name}: synthetic {cfg['m']}-sample {cfg['assay']}, hg38 "
                     f"{'chr19' if len(bins) == 1 else 'whole genome (24 chromosomes)'} @ {cfg['bin_bp']} bp "
                     f"({sum(bins.values())} bins), 2-state, per chromosome one cfixedBackgroundECM call "
                     f"iters={ECM_ITERS} (rtol 0) t_inner={T_INNER}, kappa re-weighting on, lambda off"),
        "m": cfg["m"], "n_total": int(sum(bins.values())), "chromosomes": len(bins), "bin_bp": cfg["bin_bp"],
        "longest_chromosome_bins": int(max(bins.values())), "sweeps_per_step": SWEEPS_PER_STEP,
        "l2": "inputs larger than L2: every step streams all of the genome's tracks "
              f"({2 * 4 * cfg['m'] * sum(bins.values()) / 1e9:.1f} GB) between two visits of the same bytes",
    }


# ------------------------------------------------------------------------------------------
# synthetic tracks (SURVEY 8d generator): latent bumps + slow sinusoid, per-sample offset and noise
# ------------------------------------------------------------------------------------------
def synth_host(seed: int, m: int, n: int, bin_bp: int = 25):
    rng = np.random.default_rng(seed)
    k = np.arange(n, dtype=np.float64)
    x = 0.5 * np.sin(2 * np.pi * k / 5.0e4)
    n_peaks = max(1, n // 800)
    centers = rng.integers(0, n, size=n_peaks)
    widths = rng.uniform(200 / bin_bp, 2000 / bin_bp, size=n_peaks)
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
    """Same recipe generated on the device (Philox), track by track (no [m x n] temporaries beyond the
    outputs); rows padded to ld for 16-byte aligned float4 loads."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    k = torch.arange(n, device=dev, dtype=torch.float32)
    x = 0.5 * torch.sin(2 * np.pi * k / 5.0e4)
    del k
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
    del imp, sm
    amp = 1.0 + x.abs()
    data = torch.zeros(m, ld, device=dev)
    munc = torch.ones(m, ld, device=dev)
    v0 = 0.05 + 0.25 * torch.rand(m, device=dev, generator=g)
    off = 0.05 * torch.randn(m, device=dev, generator=g)
    for j in range(m):
        row = munc[j, :n]
        torch.rand(n, device=dev, generator=g, out=row)
        row.add_(0.5).mul_(amp).mul_(v0[j])
        noise = torch.randn(n, device=dev, generator=g)
        data[j, :n] = x + off[j] + noise * row.sqrt()
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


def ecm_kwargs(data, munc):
    """Keyword arguments of one L-run call, identical for the reference and for consenrich_b200."""
    n = data.shape[1]
    return dict(matrixData=data, matrixPluginMuncInit=munc, matrixF=np.array(F_MAT, np.float32),
                matrixQ0=np.array(Q0_MAT, np.float32), intervalToBlockMap=np.zeros(n, np.int32), blockCount=1,
                stateInit=0.0, stateCovarInit=1000.0, ECM_fixedBackgroundIters=ECM_ITERS, ECM_fixedBackgroundRtol=0.0,
                pad=1e-4, ECM_robustTNu=ROBUST_NU, procPrecisionMultiplierMin=KAP_BOUNDS[0],
                procPrecisionMultiplierMax=KAP_BOUNDS[1], ECM_useObsPrecisionReweighting=False,
                ECM_useProcessPrecisionReweighting=True, ECM_useAPN=False, t_innerIters=T_INNER,
                returnIntermediates=True, returnDiagnostics=False, logIterations=False)


def cpu_ecm(mod, data, munc):
    t0 = time.perf_counter()
    out = mod.cfixedBackgroundECM(**ecm_kwargs(data, munc))
    secs = time.perf_counter() - t0
    assert out[0] == ECM_ITERS
    return secs, out


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on the host cores, same config, metric and unit.
    The hot path is single-threaded per chromosome, so "all the host threads it can use" =
    independent chromosome fits, one per thread (the loops release the GIL, cconsenrich.pyx:6578, 6740).
    A whole-genome step would take the host tens of minutes, so each step is a bounded sample: every
    thread fits one slice of a chromosome of the configuration, sized so the run ends within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    cfg = CONFIGS[args.config]
    m = cfg["m"]
    mod, kind = _cpu_module()
    cores = max(1, min(os.cpu_count() or 1, 64))
    probe = synth_host(1, m, 4_000, cfg["bin_bp"])
    secs_per_bin = min(cpu_ecm(mod, probe[0], probe[1])[0] for _ in range(2)) / 4_000
    budget = float(os.environ.get("CB200_REF_BUDGET_S", 150.0)) / max(args.steps + min(args.warmup, 1), 1)  # seconds per step
    longest = max(chrom_bins(cfg).values())
    n_sample = int(max(2_000, min(longest, budget / (3.0 * secs_per_bin))))  # 3.0: the threads share memory bandwidth
    data, munc, _ = synth_host(1729, m, n_sample, cfg["bin_bp"])

    def one(_):
        return cpu_ecm(mod, data, munc)[0]

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(min(args.warmup, 1)):
            list(ex.map(one, range(cores)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(one, range(cores)))
        dt = time.perf_counter() - t0
    value = m * n_sample * SWEEPS_PER_STEP * cores * args.steps / dt
    sample = (f"{cores} concurrent calls (one per thread) of cfixedBackgroundECM on {m} x {n_sample} bins of the "
              f"workload's matrix per step ({'oracle/_ref: the unmodified reference build' if kind == 'reference' else 'oracle port'}); "
              f"at this rate a whole-genome step takes the host {make_config(args.config)['n_total'] / (n_sample * cores) * dt / args.steps:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.config),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def _peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(m, bins, lean, log_run_of):
    """Per kernel family: (bytes per step, launches per step) this rank's chromosomes make necessary
    (DESIGN.md 4.3).  One ECM call of K = ECM_ITERS, t = T_INNER makes K t + 1 forward passes
    (K t + 1 - K plain, K - 1 with NLL that also store, 1 NLL only) and K t kappa-carrying backward passes
    plus one that publishes the smoothed tracks."""
    K, t = ECM_ITERS, T_INNER
    out = {}
    n_all = float(sum(bins))
    cells = float(m) * n_all
    calls = len(bins)
    out["fold"] = (cells * 8.0 + n_all * 32.0, calls)
    out["residuals"] = (cells * 8.0 + n_all * 8.0, calls)
    if lean:
        # per-run scan elements: 14 f64 filtering + 9 f64 smoothing per run of L bins
        runs = float(sum(n / (1 << log_run_of(n)) for n in bins))
        fwd = K * t + 1
        out["forward_compose"] = (fwd * (n_all * 20.0 + runs * 112.0), fwd * calls)
        # replay: statistics 16 (+16 with NLL) + kappa 4 in, compact track 32 out, elements in and out
        plain, nll_store, nll_only = K * t + 1 - K, K - 1, 1
        out["forward_scan"] = (plain * (n_all * 52.0 + runs * 184.0) + nll_store * (n_all * 68.0 + runs * 184.0)
                               + nll_only * (n_all * 36.0 + runs * 112.0), fwd * calls)
        out["backward_scan"] = (K * t * (n_all * 36.0 + runs * 72.0), K * t * calls)   # track 32 in, kappa 4 out
        out["backward_publish"] = (n_all * 72.0 + runs * 72.0, calls)                  # track 32 in, xs Ps lag 40 out
        out["segment_scan"] = ((2 * K * t + 1) * runs / 256.0 * 130.0, (2 * K * t + 1) * calls)      # per group of 256 runs
        out["precision_updates"] = (n_all * 16.0, 2 * calls)                           # kappa in / out of run-major order
    else:
        out["forward_scan"] = (n_all * (K * t * 76.0 + 36.0), (K * t + 1) * calls)
        out["backward_scan"] = (n_all * (K * t * 44.0 + 80.0), (K * t + 1) * calls)
    return out


def run_b200_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import consenrich_b200 as cb
    from consenrich_b200 import _lib
    from consenrich_b200.device import make_model, _p
    from consenrich_b200.sharding import assign_chromosomes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    m = cfg["m"]
    bins_all = chrom_bins(cfg)
    n_total = sum(bins_all.values())
    # chromosome sharding: longest-processing-time-first on m x bins (consenrich_b200/sharding.py)
    plan = assign_chromosomes({c: float(m) * n for c, n in bins_all.items()}, world)
    mine = plan[rank]
    my_bins = [bins_all[c] for c in mine]
    n_max = max(my_bins) if my_bins else 0

    # one explicit stream carries everything: the library's kernels, the CUDA events that time them
    # and torch's own work
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = _lib.Context(local, int(stream.cuda_stream))
    L = ctx._lib
    ld_of = lambda n: (n + 31) // 32 * 32
    # tracks: resident per chromosome when they fit the budget, otherwise the chromosomes of this rank share
    # two rotating buffers of the longest one's size (same bytes streamed, different lengths)
    need = sum(2.0 * 4 * m * ld_of(n) for n in my_bins)
    tracks = {}
    if need <= RESIDENT_BUDGET:
        for i, c in enumerate(mine):
            n = bins_all[c]
            tracks[c] = synth_device(torch, dev, 1729 + 97 * list(bins_all).index(c), m, n, ld_of(n))[:2]
        residency = f"all {len(mine)} chromosomes of the rank resident ({need / 1e9:.1f} GB)"
    elif my_bins:
        pool = [synth_device(torch, dev, 1729 + r, m, n_max, ld_of(n_max))[:2] for r in range(2)]
        for i, c in enumerate(mine):
            n = bins_all[c]
            d, v = pool[i % 2]
            # a chromosome of n bins uses the first m x ld(n) floats of the buffer as its [m][ld(n)] matrix
            tracks[c] = (d.view(-1)[: m * ld_of(n)].view(m, ld_of(n)), v.view(-1)[: m * ld_of(n)].view(m, ld_of(n)))
        residency = (f"tracks of the rank's genome ({need / 1e9:.0f} GB) exceed the resident budget: two rotating "
                     f"buffers of the longest chromosome ({2 * 2.0 * 4 * m * ld_of(n_max) / 1e9:.1f} GB) re-used per chromosome")
    model = make_model(2, F_MAT, Q0_MAT, 0.0, 1000.0, 1e-4, kap_bounds=KAP_BOUNDS, return_nll=True, use_kappa=True)
    f32 = torch.float32
    xs = torch.empty((max(n_max, 1), 2), dtype=f32, device=dev)
    Ps = torch.empty((max(n_max, 1), 2, 2), dtype=f32, device=dev)
    lag = torch.empty((max(n_max, 1), 2, 2), dtype=f32, device=dev)
    resid = torch.empty(max(n_max, 1) * m, dtype=f32, device=dev)
    kap_work = torch.ones(max(n_max, 1), dtype=f32, device=dev)
    opts = _lib.EcmOpts()
    opts.max_iters, opts.inner_iters, opts.update_lambda, opts.update_kappa, opts.want_outputs = ECM_ITERS, T_INNER, 0, 1, 1
    opts.rtol, opts.nu = 0.0, ROBUST_NU
    result = _lib.EcmResult()
    nll_sum = [0.0]

    def step():
        tot = 0.0
        for c in mine:
            n = bins_all[c]
            d, v = tracks[c]
            kap_work[:n].fill_(1.0)  # every call starts from kappa = 1, like a fresh reference call
            _lib.check(L.cb200_ecm_device(ctx.handle, C.byref(model), C.byref(opts), _p(d), _p(v), m, n, ld_of(n), None,
                                          None, _p(kap_work), _p(xs), _p(Ps), _p(lag), _p(resid), C.byref(result), None))
            assert result.iters_done == ECM_ITERS
            tot += float(result.final_nll)
        nll_sum[0] = tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----
    ctx.reset_timing()
    ctx.enable_timing(True, stride=args.timing_stride)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    kern = ctx.kernel_ms()
    kern_all = ctx.kernel_launches()
    ctx.enable_timing(False)
    nll_device = nll_sum[0]
    # diagnostic: the same K steps without the per-kernel event pairs (how much the bracketing costs)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_plain = e0.elapsed_time(e1)

    # ---- e2e through the reference-facing host API: pinned host matrices, one buffer pair of the rank's
    #      longest chromosome (filled from the device tracks), re-shaped per chromosome ----
    e2e_s, h2d, d2h, e2e_launches, nll_e2e = 0.0, 0, 0, 0, 0.0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_d = host_v = None
    if my_bins and args.e2e_steps > 0:
        c0 = max(mine, key=lambda c: bins_all[c])
        host_d = torch.empty((m, n_max), dtype=f32, pin_memory=True)
        host_v = torch.empty((m, n_max), dtype=f32, pin_memory=True)
        host_d.copy_(tracks[c0][0][:, :n_max])
        host_v.copy_(tracks[c0][1][:, :n_max])
        torch.cuda.synchronize(dev)
        flat_d, flat_v = host_d.numpy().reshape(-1), host_v.numpy().reshape(-1)

        def host_mats(n):
            return flat_d[: m * n].reshape(m, n), flat_v[: m * n].reshape(m, n)

        # Chromosomes are independent calls: a few host threads each drive their own (per-thread library
        # context = own stream + device arena), so the upload of one chromosome overlaps the sweeps of another
        # and the download of a third -- the calls themselves are the unchanged synchronous drop-in functions.
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max(1, args.e2e_threads), initializer=lambda: torch.cuda.set_device(local))

        e2e_ctxs, e2e_lock = {}, threading.Lock()

        def one(c):
            hd, hv = host_mats(bins_all[c])
            nll = float(cb.cfixedBackgroundECM(**ecm_kwargs(hd, hv))[1])
            x = _lib.default_context(local)  # this worker thread's context (launch accounting)
            with e2e_lock:
                e2e_ctxs[id(x)] = x
            return nll

        def e2e_step():
            return sum(pool.map(one, mine))

        for _ in range(2):  # warm-up: device arenas and the page-locked result pool of every worker thread
            e2e_step()
        hl0 = sum(x.launch_count for x in e2e_ctxs.values())
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nll_e2e = e2e_step()
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        e2e_launches = sum(x.launch_count for x in e2e_ctxs.values()) - hl0
        pool.shutdown()
        h2d = sum(2 * 4 * m * n for n in my_bins)                                       # tracks (kappa starts at 1 on the device)
        d2h = sum(n * (8 + 16) + (n - 1) * 16 + n * m * 4 + 4 * n + 16 for n in my_bins)  # xs, Ps, lag, residuals, kappa, scalars
    barrier()

    # ---- reduce over ranks: max time, sums of bytes / launches ----
    t = torch.tensor([ms_total, e2e_s, ms_plain], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(launches), float(h2d), float(d2h), float(e2e_launches)], dtype=torch.float64, device=dev)
    per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank[rank] = ms_total
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_s_max, ms_plain_max = float(t[0]), float(t[1]), float(t[2])
    cells = float(m) * float(n_total)
    value = cells * SWEEPS_PER_STEP * args.steps / (ms_total_max * 1e-3)
    e2e_value = cells * SWEEPS_PER_STEP * e2e_steps / e2e_s_max if e2e_s_max > 0 else None

    if rank == 0:
        peak, peak_src = _peak()
        lean = bool(int(os.environ.get("CB200_NO_LEAN", "0") or 0) == 0)
        forced = int(os.environ.get("CB200_LEAN_LOGL", "0") or 0)
        log_run_of = (lambda n: forced) if forced in (5, 6) else (lambda n: 5)
        alg = algorithmic_bytes(m, my_bins, lean, log_run_of)
        per = {k_: (v[0] / v[1]) for k_, v in kern.items() if v[1] > 0}          # ms per (sampled) launch
        tot = {k_: per[k_] * kern_all[k_] for k_ in per}                         # ms inside the timed region
        dom = max(tot, key=lambda k_: tot[k_])
        by_kernel = {}
        for k_ in per:
            if k_ in alg and alg[k_][1] > 0:
                by_kernel[k_] = (alg[k_][0] / alg[k_][1]) / (per[k_] * 1e-3) / 1e9 / peak
        alg_dom = alg[dom][0] / alg[dom][1]
        achieved = alg_dom / (per[dom] * 1e-3) / 1e9
        step_ms = ms_total_max / args.steps
        # reference-equivalent traffic of the step: the reference re-reads data + munc in every pass
        my_cells = float(m) * float(sum(my_bins))
        ref_equiv = (SWEEPS_PER_STEP * (my_cells * 8.0 + sum(my_bins) * 84.0) + ECM_ITERS * my_cells * 8.0 + my_cells * 4.0)
        traffic, traffic_src = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_full_summary.json")))
            ent = prof["kernels"][dom]
            traffic = float(ent["dram_bytes_per_launch"]) * (alg_dom / float(ent["algorithmic_bytes_per_launch"]))
            traffic_src = (f"profiles/r2_ncu_full_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of the same kernel "
                           f"captured at {ent['shape']}, scaled by this launch's algorithmic bytes")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_dom,
                    "frac_by_kernel": by_kernel,
                    "kernel_ms_per_launch": per,
                    "kernel_launches_per_step": {k_: kern_all[k_] / args.steps for k_ in per},
                    "kernel_launches_timed": {k_: kern[k_][1] for k_ in per},
                    "timing": f"CUDA events on the launching stream around every {args.timing_stride}-th launch of each "
                              "kernel family inside the timed region (an event pair costs the bracketed kernel's "
                              "neighbours their overlap, about 5 us per launch)",
                    "kernel_share_of_step": {k_: tot[k_] / ms_total for k_ in per},
                    "rank0_step": {"algorithmic_bytes": sum(v[0] for v in alg.values()),
                                   "algorithmic_GBps": sum(v[0] for v in alg.values()) / (ms_total / args.steps * 1e-3) / 1e9,
                                   "frac_of_peak": sum(v[0] for v in alg.values()) / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                                   "reference_equivalent_bytes": ref_equiv,
                                   "reference_equivalent_GBps": ref_equiv / (ms_total / args.steps * 1e-3) / 1e9,
                                   "note": "algorithmic = what the kernels of this design have to move, summed over the step; "
                                           "reference-equivalent = the reference re-reads data+munc in every pass (here they are "
                                           "folded once per call), so that figure may exceed the HBM peak and is not a roofline fraction"}}
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": None, "sample": None}
        check = {"final_nll_sum_device": nll_device, "final_nll_sum_e2e": nll_e2e,
                 "note": "device and e2e legs run different synthetic bytes; parity is the block below"}
        if world == 1 and not args.no_cpu_baseline and my_bins and host_d is not None:
            cpu, chk = cpu_baseline_and_check(cb, m, host_d.numpy(), host_v.numpy(), args.cpu_bins)
            check.update(chk)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.config),
            "clocks": clocks.summary(), "gpu_launches": int(cnt[0]),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(cnt[1]), "d2h_bytes_per_step": int(cnt[2]),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s_max / e2e_steps, "gpu_launches": int(cnt[3]),
                    "api": "per chromosome consenrich_b200.cfixedBackgroundECM -> cb200_host_ecm (pinned host matrices in, "
                           "page-locked results out), chromosomes spread over "
                           f"{args.e2e_threads} host threads (one library context each) so that copies and sweeps of different "
                           "chromosomes overlap; host matrices: one pinned pair of the rank's longest chromosome, re-shaped per "
                           "chromosome",
                    "limiter": "the bus: every step moves h2d + d2h bytes between host memory and the GPUs (68 GB/s with one "
                               "GPU, both directions overlapped); with N ranks the same bytes share the host's memory and "
                               "PCIe complex, so this number scales far below the device-resident one"},
            "roofline": roofline, "cpu_baseline": cpu,
            "placement": {"sharding": "chromosomes bin-packed onto ranks, longest first; no collective on the data path",
                          "chromosomes_per_rank": [len(p) for p in plan],
                          "bins_per_rank": [int(sum(bins_all[c] for c in p)) for p in plan],
                          "ms_per_step_by_rank": [float(x) / args.steps for x in per_rank.tolist()],
                          "imbalance": (max(sum(bins_all[c] for c in p) for p in plan) * world / float(n_total)),
                          "limiter": "the rank holding the most bins (chr1 is 8 % of the genome); no communication",
                          "residency_rank0": residency if my_bins else "no chromosome assigned",
                          "lean_sweeps": lean},
            "check": check,
            "ms_per_step_without_kernel_events": ms_plain_max / args.steps,
            "value_without_kernel_events": cells * SWEEPS_PER_STEP * args.steps / (ms_plain_max * 1e-3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_and_check(cb, m, host_d, host_v, ns):
    """The reference's own cfixedBackgroundECM on the first `ns` bins of the longest chromosome's host
    matrices (1 core), and the GPU call on the SAME slice compared output by output."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_util import max_violation
    mod, kind = _cpu_module()
    ns = int(min(ns, host_d.shape[1]))
    d = np.ascontiguousarray(host_d[:, :ns])
    v = np.ascontiguousarray(host_v[:, :ns])
    secs, ref = cpu_ecm(mod, d, v)
    got = cb.cfixedBackgroundECM(**ecm_kwargs(d, v))
    cpu = {"value": m * ns * SWEEPS_PER_STEP / secs, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"the first {ns} bins of the longest chromosome's matrices ({m} x {ns}), one cfixedBackgroundECM "
                     f"call ({secs:.2f} s)"}

    def err_over_scale(a, b, comp=False):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if comp:
            sc = np.max(np.abs(b.reshape(len(b), -1)), axis=0, keepdims=True)
            return float(np.max(np.abs(a.reshape(len(a), -1) - b.reshape(len(b), -1)) / np.maximum(sc, 1e-300)))
        return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))

    viol = {"state": max_violation(got[2], ref[2]), "covariance": max_violation(got[3], ref[3], "component"),
            "lag_covariance": max_violation(got[4], ref[4], "component"), "residuals": max_violation(got[5], ref[5]),
            "kappa": max_violation(got[7], ref[7], rtol=2e-4)}
    chk = {"slice": f"{m} x {ns}", "against": kind, "iters": [int(got[0]), int(ref[0])],
           "nll_rel_diff": abs(got[1] - ref[1]) / max(abs(ref[1]), 1.0),
           "max_err_over_scale": {"state": err_over_scale(got[2], ref[2]), "covariance": err_over_scale(got[3], ref[3], True),
                                  "residuals": err_over_scale(got[5], ref[5]), "kappa": err_over_scale(got[7], ref[7])},
           "max_err_over_tolerance": viol,
           "tolerance": "tests/parity_util.py: |got - want| <= 1e-4 |want| + 1e-5 scale (kappa: 2e-4 |want|)"}
    ok = got[0] == ref[0] and chk["nll_rel_diff"] <= 1e-7 and all(x <= 1.0 for x in viol.values())
    chk["within_tolerance"] = bool(ok)
    if not ok:
        raise SystemExit(f"bench.py: GPU result differs from the reference beyond the stated tolerance: {json.dumps(chk)}")
    return cpu, chk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2, help="0 skips the e2e leg (profiling runs)")
    ap.add_argument("--e2e-threads", type=int, default=3, help="host threads that drive chromosomes concurrently in the e2e leg")
    ap.add_argument("--timing-stride", type=int, default=7,
                    help="every n-th launch of a kernel family is event-timed (7 is coprime to the 24 chromosomes and the 16 / 31 passes of a call, so the samples cover every position)")
    ap.add_argument("--cpu-bins", type=int, default=None, help="bins of the cpu_baseline / check slice")
    args = ap.parse_args()
    if args.cpu_bins is None:
        args.cpu_bins = max(50_000, int(2.0e8 / CONFIGS[args.config]["m"]))  # ~15-25 s on one core
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
        args.steps = 10 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else max(3, args.warmup)
        run_b200_arm(args)


if __name__ == "__main__":
    main()
