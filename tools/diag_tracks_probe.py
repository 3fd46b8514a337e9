"""Times core._perIntervalOutputDiagnosticTracks (the reference's per-interval Python loop) against the
device hook (diagnostics; needs oracle/_ref/driver).  usage: python tools/diag_tracks_probe.py [n] [m]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "driver"))
import consenrich.core as core  # noqa: E402
import consenrich_b200 as cb  # noqa: E402
from test_driver_hooks import diag_inputs, diag_state_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kw = diag_inputs(np.random.default_rng(0), m, n, 2, "kappa")
kw["stateModel"] = diag_state_model(core, 2)
t0 = time.perf_counter()
want = core._perIntervalOutputDiagnosticTracks(**kw)
t_ref = time.perf_counter() - t0
cb.install_driver(core)
try:
    core._perIntervalOutputDiagnosticTracks(**kw)
    t0 = time.perf_counter()
    got = core._perIntervalOutputDiagnosticTracks(**kw)
    t_dev = time.perf_counter() - t0
finally:
    cb.uninstall_driver(core)
worst = max(float(np.max(np.abs(got[k].astype(np.float64) - want[k]) / np.maximum(np.abs(want[k]), 1e-30))) for k in want)
print(json.dumps({"what": "core._perIntervalOutputDiagnosticTracks", "tracks": m, "intervals": n, "seconds_reference": t_ref,
                  "seconds_with_driver_hook": t_dev, "speedup": t_ref / t_dev, "max_relative_difference": worst}))
