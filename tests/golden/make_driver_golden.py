"""tests/golden/make_driver_golden.py -- regenerates tests/golden/driver_golden.npz.

Runs the UNMODIFIED reference driver functions (``oracle/_ref/driver``: the reference's ``core.py``) on small
seeded inputs for the driver-side reductions that have a device version (core.py:2647-2700).

    python tests/golden/make_driver_golden.py      # needs oracle/_ref/driver (`make -C oracle ref`)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))


def sign_change_inputs(rng, m, n):
    from conftest import synth_tracks
    data, munc = synth_tracks(int(rng.integers(1 << 30)), m, n, masked_frac=0.05)
    data[rng.random((m, n)) < 0.01] = np.nan
    state = np.nan_to_num(data[0] * 0.7 + rng.normal(0, 0.05, n)).astype(np.float32)
    if n > 10:
        state[5] = np.inf
    background = rng.normal(0, 0.1, n).astype(np.float32)
    return state, data, munc, background


def main():
    drv = os.path.join(ROOT, "oracle", "_ref", "driver")
    if not os.path.isdir(os.path.join(drv, "consenrich")):
        raise SystemExit("oracle/_ref/driver missing: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, drv)
    import consenrich.core as core
    rng = np.random.default_rng(20261021)
    out = {}
    for name, m, n, with_bg in (("m1_n1", 1, 1, False), ("m3_n50", 3, 50, True), ("m6_n5000", 6, 5000, True),
                                ("m4_n3001", 4, 3001, False)):
        state, data, munc, bg = sign_change_inputs(rng, m, n)
        val = core._relativeSignChangePerKB(state, data, munc, intervalSizeBP=25, background=bg if with_bg else None,
                                            pad=1e-4)
        out[f"{name}/state"], out[f"{name}/data"], out[f"{name}/munc"] = state, data, munc
        if with_bg:
            out[f"{name}/background"] = bg
        out[f"{name}/value"] = np.float64(np.nan if val is None else val)
    path = os.path.join(HERE, "driver_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
