import cProfile, pstats, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import consenrich_b200 as cb
m, n = bench.M_TRACKS, bench.N_BINS
data, munc, kap = bench.synth_host(1, m, n)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
data, munc = pin(data), pin(munc)
kw = bench.ecm_kwargs(data, munc)
for _ in range(3): r = cb.cfixedBackgroundECM(**kw)
t0 = time.perf_counter()
for _ in range(5): r = cb.cfixedBackgroundECM(**kw)
print("ms per call", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): r = cb.cfixedBackgroundECM(**kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
# raw copies
d = torch.empty(m * n, dtype=torch.float32, device="cuda")
h = torch.empty(m * n, dtype=torch.float32).pin_memory()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, "94 MB pinned:", round(dt * 1e3, 2), "ms", round(m * n * 4 / dt / 1e9, 1), "GB/s")
