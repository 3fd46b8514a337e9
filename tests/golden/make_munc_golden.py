"""tests/golden/make_munc_golden.py -- regenerates tests/golden/munc_golden.npz.

Runs the UNMODIFIED reference (``oracle/_ref``) on small seeded inputs for the dense kernels of the
observation-noise stage (cconsenrich.pyx:5547-5740) and stores inputs + outputs.

    python tests/golden/make_munc_golden.py      # needs oracle/_ref (`make -C oracle ref`)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def local_evidence(rng, m, n):
    """Positive, heavy-tailed per-cell evidence (squared residual scale), float32."""
    k = np.arange(n)
    base = 0.2 * (1.5 + np.sin(k / 90.0))[None, :] * rng.uniform(0.5, 2.0, (m, 1))
    return np.ascontiguousarray((base * rng.chisquare(1.0, (m, n)) + 1e-6).astype(np.float32))


def exclude_mask(rng, m, n, mode):
    if mode == 0:
        return None
    shape = (n,) if mode == 1 else (m, n)
    mask = (rng.random(shape) < 0.08).astype(np.uint8)
    if n > 60:  # a blacklisted stretch longer than any test window
        mask[..., n // 3: n // 3 + 50] = 1
    return mask


SEED_POSITIONAL = ("matrixData", "matrixMunc", "stateMean", "stateVariance")
SEED_OUTPUTS = ("moment", "rhoOut", "omegaRaw", "omegaOut", "local", "variance")


def seed_case(rng, m, n, variant):
    """Inputs of one cMuncObservationMomentSeedPass call; `variant` picks the branch of pyx:4843-5040."""
    c = dict(matrixData=rng.normal(0, 1, (m, n)).astype(np.float32),
             matrixMunc=rng.uniform(0.05, 1.0, (m, n)).astype(np.float32),
             stateMean=rng.normal(0, 0.3, n).astype(np.float32), stateVariance=rng.uniform(0, 0.2, n).astype(np.float32))
    if variant in ("update", "fixed", "gaussian"):
        c["background"] = rng.normal(0, 0.1, n).astype(np.float32)
        c["gVariance"] = rng.uniform(-0.01, 0.05, n).astype(np.float32)
        c["countFloor"] = rng.uniform(0, 0.3, (m, n)).astype(np.float32)
    if variant == "update":
        c["omegaIn"] = rng.uniform(0.2, 3.0, n).astype(np.float32)
        c["activeMask"] = (rng.random(n) < 0.9).astype(np.uint8)
        c.update(pad=0.01, studentTdf=4.0, dOmega=6.0, omegaMin=0.5, omegaMax=1.5, varianceFloor=1e-3, varianceCap=1.5)
    elif variant == "fixed":
        c["omegaIn"] = rng.uniform(0.2, 3.0, n).astype(np.float32)
        c["rhoIn"] = rng.uniform(0.2, 2.0, (m, n)).astype(np.float32)
        c["activeMask"] = (rng.random((m, n)) < 0.85).astype(np.uint8)
        c["updateWeights"] = False
    elif variant == "unweighted":
        c["useSeedWeights"] = False
        c["activeMask"] = (rng.random((m, n)) < 0.7).astype(np.uint8)
        c["countFloor"] = rng.uniform(0, 0.3, (m, n)).astype(np.float32)
        c["varianceCap"] = 0.8
    elif variant == "gaussian":
        c.update(studentT=False, varianceFloor=0.05, varianceCap=0.6)
    return c


def main():
    ref = O.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261020)
    out = {}
    # the reference's own known-answer case (tests/test_core.py:1535-1556)
    le = np.asarray([[1.0, 100.0, 3.0, 5.0, 7.0], [2.0, 4.0, 8.0, 16.0, 32.0]], np.float32)
    mk = np.asarray([0, 1, 0, 0, 0], np.uint8)
    out["smooth/ref_case/local"], out["smooth/ref_case/mask"] = le, mk
    out["smooth/ref_case/window"], out["smooth/ref_case/eps"] = np.int64(3), np.float64(1.0e-4)
    out["smooth/ref_case/out"] = ref.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=mk, eps=1.0e-4)
    specs = [("m1_n1_w1", 1, 1, 1, 0), ("m2_n7_w4", 2, 7, 4, 1), ("m3_n50_w64", 3, 50, 64, 0),
             ("m4_n1000_w9", 4, 1000, 9, 2), ("m6_n5000_w40", 6, 5000, 40, 1), ("m2_n4099_w2048", 2, 4099, 2048, 2)]
    for name, m, n, w, mode in specs:
        le = local_evidence(rng, m, n)
        mk = exclude_mask(rng, m, n, mode)
        eps = 1.0e-4 if mode else 1.0e-12
        res = ref.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=eps)
        out[f"smooth/{name}/local"] = le
        if mk is not None:
            out[f"smooth/{name}/mask"] = mk
        out[f"smooth/{name}/window"], out[f"smooth/{name}/eps"] = np.int64(w), np.float64(eps)
        out[f"smooth/{name}/out"] = res
    # cFinalizeMuncEBTrack: the reference's own known-answer case (tests/test_core.py:1481-1522) + seeded ones
    fin = [("ref_case", np.asarray([0.05, 0.20, 4.00, 0.001], np.float32), np.asarray([0.35, 0.60, 10.00, 0.02], np.float32),
            np.asarray([np.nan, 0.25, 2.00, 0.00], np.float32), dict(nuLocal=2.0, nuPrior=3.0, useEB=True,
                                                                       varianceFloor=0.01, varianceCap=5.0)),
           ("ref_case_no_eb", np.asarray([0.05, 0.20, 4.00, 0.001], np.float32), None,
            np.asarray([np.nan, 0.25, 2.00, 0.00], np.float32), dict(useEB=False, varianceFloor=0.01, varianceCap=5.0))]
    for name, n, with_cf, opts in (("n1", 1, True, dict(nuLocal=4.0, nuPrior=9.5, varianceFloor=1e-4, varianceCap=50.0)),
                                   ("n5000", 5000, True, dict(nuLocal=37.0, nuPrior=12.25, varianceFloor=1e-3,
                                                              varianceCap=3.0)),
                                   ("n3001_defaults", 3001, False, dict(nuLocal=6.0, nuPrior=2.0))):
        loc = local_evidence(rng, 1, n)[0]
        pri = (loc * rng.uniform(0.3, 3.0, n)).astype(np.float32)
        cf = None
        if with_cf:
            cf = rng.uniform(0.0, 0.5, n).astype(np.float32)
            cf[rng.random(n) < 0.2] = np.nan
            cf[rng.random(n) < 0.2] = 0.0
        fin.append((name, loc, pri, cf, opts))
    for name, loc, pri, cf, opts in fin:
        res, diag = ref.cFinalizeMuncEBTrack(loc, priorVarianceTrack=pri, countFloor=cf, **opts)
        out[f"finalize/{name}/local"] = loc
        if pri is not None:
            out[f"finalize/{name}/prior"] = pri
        if cf is not None:
            out[f"finalize/{name}/countFloor"] = cf
        for k_, v_ in opts.items():
            out[f"finalize/{name}/opt_{k_}"] = np.asarray(v_)
        out[f"finalize/{name}/out"] = res
        for k_, v_ in diag.items():
            out[f"finalize/{name}/diag_{k_}"] = np.asarray(v_)
    # cMuncObservationMomentSeedPass: every branch of pyx:4843-5040
    for name, m, n, variant in (("weighted_update", 4, 600, "update"), ("weighted_fixed", 3, 501, "fixed"),
                                ("unweighted_masked", 5, 400, "unweighted"), ("gaussian_clipped", 2, 300, "gaussian"),
                                ("single_cell", 1, 1, "update")):
        c = seed_case(rng, m, n, variant)
        res = ref.cMuncObservationMomentSeedPass(c["matrixData"], c["matrixMunc"], c["stateMean"], c["stateVariance"],
                                                 **{k: v for k, v in c.items() if k not in SEED_POSITIONAL})
        for k, v in c.items():
            out[f"seed/{name}/in_{k}"] = np.asarray(v)
        for k, v in zip(SEED_OUTPUTS, res):
            out[f"seed/{name}/out_{k}"] = v
    # cEMA: the reference's own known-answer input (tests/test_core.py:1357-1369) in both types + seeded tracks
    ema = [("ref_case_f32", np.array([0.0, 2.0, -1.0, 5.0, 4.0, 7.0], np.float32), 0.35),
           ("ref_case_f64", np.array([0.0, 2.0, -1.0, 5.0, 4.0, 7.0], np.float64), 0.35)]
    for name, n, dt, alpha in (("n1_f32", 1, np.float32, 0.5), ("n3000_f32", 3000, np.float32, 2.0 / 42.0),
                               ("n3001_f64", 3001, np.float64, 0.1), ("n1025_f32_tiny", 1025, np.float32, 1e-3),
                               ("n700_f64_one", 700, np.float64, 1.0)):
        ema.append((name, (0.3 + np.abs(rng.normal(size=n))).astype(dt), alpha))
    for name, x, alpha in ema:
        out[f"ema/{name}/x"], out[f"ema/{name}/alpha"] = x, np.float64(alpha)
        out[f"ema/{name}/out"] = ref.cEMA(x, alpha)
    path = os.path.join(HERE, "munc_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
