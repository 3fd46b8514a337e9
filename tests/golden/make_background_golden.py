"""tests/golden/make_background_golden.py -- regenerates tests/golden/background_golden.npz.

Runs the UNMODIFIED reference (``oracle/_ref``) on small seeded inputs for the background-track
functions (cconsenrich.pyx:944-1096, 9675-9724) and stores inputs + outputs.

    python tests/golden/make_background_golden.py      # needs oracle/_ref (`make -C oracle ref`)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def background_inputs(rng, n, zero_frac=0.05):
    """Weight / rhs tracks of the kind core.py:5064-5083 forms: sums of inverse variances, and a smooth
    background plus noise weighted by them; stretches of zero weight where every sample is masked."""
    k = np.arange(n)
    w = rng.uniform(5.0, 200.0, n)
    gaps = rng.random(n) < zero_frac
    w[gaps] = 0.0
    if n > 40:
        s = int(rng.integers(5, n - 30))
        w[s:s + 20] = 0.0  # a masked stretch
    if n >= 2:
        w[0] = max(w[0], 1.0)
        w[-1] = max(w[-1], 1.0)
    g = 0.4 * np.sin(2 * np.pi * k / max(n, 16) * 2.0) + 0.1 * np.cos(k / 7.0)
    rhs = w * (g + rng.normal(0, 0.2, n))
    return w, rhs


def main():
    ref = O.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261019)
    out = {}
    specs = [  # name, n, lam, lamFirst, zeroCenter
        ("n2", 2, 8.0, 0.5, True), ("n3", 3, 8.0, 0.0, False), ("n4", 4, 2.0, 1.0, True), ("n5_odd", 5, 128.0, 0.0, True),
        ("n257_second_only", 257, 128.0, 0.0, True), ("n1000_both", 1000, 64.0, 4.0, False),
        ("n1501_first_only", 1501, 0.0, 16.0, True), ("n4096_stiff", 4096, 1.0e4, 0.0, True),
        ("n3000_no_penalty", 3000, 0.0, 0.0, False),
    ]
    for name, n, lam, lam1, zc in specs:
        w, rhs = background_inputs(rng, n, 0.0 if name == "n3000_no_penalty" else 0.05)
        if name == "n3000_no_penalty":
            w = np.maximum(w, 1.0)
        x = ref.csolveZeroCenteredBackground(w, rhs, lam, zc, lamFirst=lam1)
        out[f"solve/{name}/weight"], out[f"solve/{name}/rhs"], out[f"solve/{name}/out"] = w, rhs, x
        out[f"solve/{name}/lam"], out[f"solve/{name}/lamFirst"] = np.float64(lam), np.float64(lam1)
        out[f"solve/{name}/zeroCenter"] = np.bool_(zc)
    for name, m, n in (("m3_n100", 3, 100), ("m10_n2049", 10, 2049), ("m1_n7", 1, 7)):
        res = rng.normal(0, 1.0, (m, n)).astype(np.float32)
        inv = (1.0 / rng.uniform(0.05, 2.0, (m, n))).astype(np.float32)
        inv[:, rng.random(n) < 0.1] = 0.0
        w, r, sup = ref.cbackgroundWeightedStatsWithSupport(res, inv)
        out[f"stats/{name}/resid"], out[f"stats/{name}/inv"] = res, inv
        out[f"stats/{name}/weight"], out[f"stats/{name}/rhs"], out[f"stats/{name}/support"] = w, r, np.int64(sup)
    path = os.path.join(HERE, "background_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
