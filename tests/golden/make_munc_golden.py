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
    path = os.path.join(HERE, "munc_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
