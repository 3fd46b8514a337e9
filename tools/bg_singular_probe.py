"""Raise / no-raise behaviour of csolveZeroCenteredBackground on singular and near-singular systems: device vs oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import consenrich_b200 as cb
from oracle import oracle as O
O.build()
rng = np.random.default_rng(0)
n = 5000
cases = {
    "all-zero weights, lamFirst 0": (np.zeros(n), rng.normal(size=n), 4.0, 0.0),
    "all-zero weights, lamFirst 1": (np.zeros(n), rng.normal(size=n), 4.0, 1.0),
    "one positive weight, lamFirst 0": (np.eye(1, n, 7)[0] * 2.0, rng.normal(size=n), 4.0, 0.0),
    "two positive weights, lamFirst 0": (np.eye(1, n, 7)[0] * 2.0 + np.eye(1, n, 4000)[0], rng.normal(size=n), 4.0, 0.0),
    "tiny weights 1e-14": (np.full(n, 1e-14), rng.normal(size=n), 4.0, 0.0),
    "tiny weights 1e-10": (np.full(n, 1e-10), rng.normal(size=n), 4.0, 0.0),
    "lam 0, zero weight inside": (np.where(np.arange(n) == 100, 0.0, 1.0), rng.normal(size=n), 0.0, 0.0),
    "lam 0, tiny weight inside": (np.where(np.arange(n) == 100, 1e-13, 1.0), rng.normal(size=n), 0.0, 0.0),
    "regular": (rng.uniform(0.5, 2, n), rng.normal(size=n), 128.0, 0.0),
}
for name, (w, r, lam, lam1) in cases.items():
    for zc in (True, False):
        out = []
        for mod in (cb, O):
            try:
                x = mod.csolveZeroCenteredBackground(w, r, lam, zc, lamFirst=lam1)
                out.append(("ok", float(np.abs(x).max()), bool(np.all(np.isfinite(x)))))
            except Exception as e:
                out.append((type(e).__name__, str(e)[:70]))
        print(f"{name:38s} zeroCenter={zc!s:5s} device={out[0]} oracle={out[1]}", flush=True)
