"""The reference's own runConsenrich driver, on its Cython kernels and again with the B200 kernels
installed: wall time of each and where the installed run still spends it (diagnostics; needs
oracle/_ref/driver, built by oracle/build_ref_driver.sh).

usage: python tools/run_consenrich_e2e.py [n_intervals] [tracks]"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DRV = os.path.join(ROOT, "oracle", "_ref", "driver")
if not os.path.isdir(os.path.join(DRV, "consenrich")):
    raise SystemExit("oracle/_ref/driver missing: run oracle/build_ref_driver.sh where /root/reference exists")
sys.path.insert(0, DRV)
import consenrich.core as core  # noqa: E402
import consenrich_b200 as cb  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_run_consenrich_gpu import BASE, CASES, _tracks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
data, munc = _tracks(7, m, n)
kw = {**BASE, **CASES["cli_defaults"]}


def run():
    t0 = time.perf_counter()
    out = core.runConsenrich(data, munc, **kw)
    return time.perf_counter() - t0, out


t_ref, want = run()
cb.install()
try:
    run()  # warm-up: context, buffers, pinned pool
    ctx = cb._lib.default_context()
    l0 = ctx.launch_count
    t_gpu, got = run()
    launches = ctx.launch_count - l0
    cb.install_driver(core)
    try:
        run()
        t_gpu_driver, got_driver = run()
        pr = cProfile.Profile()
        pr.enable()
        run()
        pr.disable()
    finally:
        cb.uninstall_driver(core)
finally:
    cb.uninstall()
err = float(np.abs(got[0].astype(np.float64) - want[0]).max() / np.abs(want[0]).max())
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(14)
print(json.dumps({"what": "core.runConsenrich, CLI defaults (fitBackground on), synthetic tracks", "tracks": m, "intervals": n,
                  "seconds_reference_kernels": t_ref, "seconds_b200_kernels": t_gpu, "speedup": t_ref / t_gpu,
                  "seconds_b200_kernels_and_driver_hooks": t_gpu_driver, "speedup_with_driver_hooks": t_ref / t_gpu_driver,
                  "driver_hooks_change_the_result": bool(not np.array_equal(got_driver[0], got[0])),
                  "kernel_launches": int(launches), "max_state_err_over_scale": err,
                  "final_nll": [float(want[-1]["final_nll"]), float(got[-1]["final_nll"])]}))
print(buf.getvalue())
