"""tests/golden/make_golden.py -- regenerates tests/golden/*.npz.

Runs the UNMODIFIED reference (``oracle/_ref``: ``cconsenrich.pyx`` compiled by
``oracle/build_ref.sh``) on small seeded inputs and stores inputs + outputs.  The fixtures
are what pins the oracle (and, through it, the CUDA path) on machines where
``/root/reference`` and ``oracle/_ref`` are absent.

    python tests/golden/make_golden.py      # needs oracle/_ref (run `make -C oracle ref` first)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth(rng, m, n, masked_frac=0.0):
    """Small cousin of the benchmark generator: latent bumps + per-sample noise."""
    k = np.arange(n)
    x = 0.3 * np.sin(2 * np.pi * k / max(n, 8) * 3.0)
    for _ in range(max(1, n // 200)):
        c, w, h = rng.integers(0, n), rng.uniform(3, 30), rng.uniform(0.5, 4.0)
        x = x + h * np.exp(-0.5 * ((k - c) / w) ** 2)
    v0 = rng.uniform(0.05, 0.3, size=(m, 1))
    munc = (v0 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n))).astype(np.float32)
    data = (x[None, :] + rng.normal(0, 0.05, size=(m, 1)) + rng.normal(size=(m, n)) * np.sqrt(munc)).astype(np.float32)
    if masked_frac > 0:
        mask = rng.random((m, n)) < masked_frac
        munc[mask] = np.float32(1.0e30)  # constants.py:387 masked-observation variance
    return data, munc


def main():
    ref = O.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261018)
    F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
    cases = {}
    specs = [
        # name, m, n, masked, variant
        ("plain_m3_n257", 3, 257, 0.0, "plain"),
        ("weights_m10_n1500", 10, 1500, 0.05, "weights"),
        ("cli_bounds_m5_n900", 5, 900, 0.0, "cli"),
        ("apn_m4_n400", 4, 400, 0.0, "apn"),
        ("tiny_m2_n1", 2, 1, 0.0, "plain"),
        ("tiny_m2_n3", 2, 3, 0.0, "weights"),
    ]
    for name, m, n, masked, variant in specs:
        data, munc = synth(rng, m, n, masked)
        Q0 = np.array([[rng.uniform(1e-4, 5e-3), 0.0], [0.0, rng.uniform(1e-5, 1e-3)]], np.float32)
        bm = (np.arange(n) // 64).astype(np.int32)
        kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=bm,
                  blockCount=int(bm.max()) + 1, stateInit=float(np.float32(data[:, 0].mean())),
                  stateCovarInit=1000.0, pad=1.0e-4, returnNLL=True)
        extra = {}
        if variant in ("weights", "cli"):
            extra["lambdaExp"] = (0.1 + 5.0 * rng.random(n)).astype(np.float32)
            extra["processPrecExp"] = np.exp(rng.normal(0, 2.0, n)).astype(np.float32)
            qs = (0.5 + rng.random(n)).astype(np.float32)
            qs[0] = 1.0
            extra["processQScale"] = qs
        if variant == "cli":
            extra.update(procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3, storeNLLInD=True)
        if variant == "apn":
            extra.update(ECM_useAPN=True)
        for dim in (2, 1):
            st = dict(stateForward=np.empty((n, dim), np.float32),
                      stateCovarForward=np.empty((n, dim, dim), np.float32),
                      pNoiseForward=np.zeros((n, dim, dim), np.float32), vectorD=np.empty(n, np.float32))
            if dim == 2:
                r = ref.cforwardPass(matrixF=F, **kw, **extra, **st)
                b = ref.cbackwardPass(matrixData=data, matrixF=F, stateForward=st["stateForward"],
                                      stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"])
            else:
                r = ref.cforwardPassLevel(**kw, **extra, **st)
                b = ref.cbackwardPassLevel(matrixData=data, stateForward=st["stateForward"],
                                           stateCovarForward=st["stateCovarForward"],
                                           pNoiseForward=st["pNoiseForward"])
            pre = f"{name}/d{dim}/"
            cases[pre + "phiHat"] = np.float32(r[0])
            cases[pre + "sumNLL"] = np.float64(r[3])
            cases[pre + "vectorD"] = st["vectorD"]
            cases[pre + "stateForward"] = st["stateForward"]
            cases[pre + "stateCovarForward"] = st["stateCovarForward"]
            cases[pre + "pNoiseForward"] = st["pNoiseForward"]
            for nm, arr in zip(("stateSmoothed", "stateCovarSmoothed", "lagCovSmoothed", "postFitResiduals"), b):
                cases[pre + nm] = arr
        cases[f"{name}/data"] = data
        cases[f"{name}/munc"] = munc
        cases[f"{name}/Q0"] = Q0
        cases[f"{name}/blockMap"] = bm
        cases[f"{name}/stateInit"] = np.float32(kw["stateInit"])
        cases[f"{name}/variant"] = np.array(variant)
        for k_, v_ in extra.items():
            cases[f"{name}/extra/{k_}"] = np.asarray(v_)
    np.savez_compressed(os.path.join(HERE, "sweep_golden.npz"), **cases)

    # ECM goldens: fixed iteration budget (rtol=0) and free-running.
    ecm = {}
    ecm_specs = [
        ("kappa_only_m4_n600", 4, 600, dict(ECM_useObsPrecisionReweighting=False, ECM_fixedBackgroundIters=3,
                                           ECM_fixedBackgroundRtol=0.0, procPrecisionMultiplierMin=5e-3,
                                           procPrecisionMultiplierMax=5e3)),
        ("both_m6_n800", 6, 800, dict(ECM_fixedBackgroundIters=4, ECM_fixedBackgroundRtol=0.0, t_innerIters=3)),
        ("free_m3_n500", 3, 500, dict(ECM_fixedBackgroundIters=25, ECM_fixedBackgroundRtol=1e-4)),
        ("tiny_m2_n4", 2, 4, dict(ECM_fixedBackgroundIters=3)),
    ]
    for name, m, n, opts in ecm_specs:
        data, munc = synth(rng, m, n, 0.02 if n > 100 else 0.0)
        Q0 = np.array([[2e-3, 0.0], [0.0, 4e-4]], np.float32)
        bm = np.zeros(n, np.int32)
        kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=bm, blockCount=1,
                  stateInit=0.0, stateCovarInit=1000.0, returnIntermediates=True, returnDiagnostics=True,
                  logIterations=False, **opts)
        for dim in (2, 1):
            out = ref.cfixedBackgroundECM(matrixF=F, **kw) if dim == 2 else ref.cfixedBackgroundECMLevel(**kw)
            pre = f"{name}/d{dim}/"
            ecm[pre + "itersDone"] = np.int64(out[0])
            ecm[pre + "nll"] = np.float64(out[1])
            for nm, arr in zip(("stateSmoothed", "stateCovarSmoothed", "lagCovSmoothed", "postFitResiduals",
                                "lambdaExp", "processPrecExp"), out[2:8]):
                if arr is not None:
                    ecm[pre + nm] = arr
            ecm[pre + "converged"] = np.bool_(out[8]["converged"])
        ecm[f"{name}/data"] = data
        ecm[f"{name}/munc"] = munc
        ecm[f"{name}/Q0"] = Q0
        for k_, v_ in opts.items():
            ecm[f"{name}/opts/{k_}"] = np.asarray(v_)
    np.savez_compressed(os.path.join(HERE, "ecm_golden.npz"), **ecm)
    for f in ("sweep_golden.npz", "ecm_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
