"""Measured GPU-vs-oracle error of every cfixedBackgroundECM parity case in tests/test_gpu_parity.py and
tests/test_lean_sweeps.py (max |err| / scale per output): the stated tolerances are set from these."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import consenrich_b200 as cb
from oracle import oracle as O
from conftest import synth_tracks
import test_gpu_parity as T

O.build()


def err(a, b, comp=False):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    a2, b2 = a.reshape(len(a), -1), b.reshape(len(b), -1)
    sc = np.max(np.abs(b2), axis=0, keepdims=True) if comp else np.max(np.abs(b2))
    return float(np.max(np.abs(a2 - b2) / np.maximum(sc, 1e-300)))


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def report(label, a, b):
    out = {"nll_rel": abs(a[1] - b[1]) / max(abs(b[1]), 1.0), "state": err(a[2], b[2]), "P": err(a[3], b[3], True),
           "lag": err(a[4], b[4], True), "resid": err(a[5], b[5])}
    for i, nm in ((6, "lambda"), (7, "kappa")):
        if a[i] is not None:
            out[nm] = err(a[i], b[i])
            out[nm + "_rel"] = relerr(a[i], b[i])
    print(label, "iters", a[0], b[0], {k: f"{v:.2e}" for k, v in out.items()}, flush=True)


fixed = [
    dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2),
    dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, t_innerIters=3,
         ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3),
    dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, ECM_useProcessPrecisionReweighting=False),
    dict(ECM_fixedBackgroundIters=4, ECM_fixedBackgroundRtol=0.0, t_innerIters=1, ECM_robustTNu=4.0),
]
data, munc = synth_tracks(909, 8, 6000, masked_frac=0.02)
for dim in (2, 1):
    for i, o in enumerate(fixed):
        report(f"fixed d{dim} #{i}", T._ecm(cb, dim, data, munc, **o), T._ecm(O, dim, data, munc, **o))
data, munc = synth_tracks(31, 5, 3000)
rng = np.random.default_rng(5)
qs = (0.5 + rng.random(3000)).astype(np.float32); qs[0] = 1.0
o = dict(ECM_fixedBackgroundIters=25, ECM_fixedBackgroundRtol=1e-4, processQScale=qs,
         lambdaExpInit=(0.2 + 6 * rng.random(3000)).astype(np.float32),
         processPrecExpInit=np.exp(rng.normal(0, 1, 3000)).astype(np.float32))
for dim in (2, 1):
    report(f"free d{dim}", T._ecm(cb, dim, data, munc, **o), T._ecm(O, dim, data, munc, **o))
for m, n in ((7, 40_003), (3, 16385), (2, 900_001)):
    data, munc = synth_tracks(4000 + n, m, n, masked_frac=0.03)
    for o in (dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2,
                   ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3),
              dict(ECM_fixedBackgroundIters=12, ECM_fixedBackgroundRtol=1e-3, t_innerIters=2)):
        report(f"{m}x{n} {'cli' if 'ECM_useObsPrecisionReweighting' in o else 'lam+kap free'}",
               T._ecm(cb, 2, data, munc, **o), T._ecm(O, 2, data, munc, **o))
