"""Product-level parity: the reference's OWN driver, `consenrich.core.runConsenrich`
(installed under oracle/_ref/driver by oracle/build_ref_driver.sh), run twice on the same inputs --
once on its Cython kernels, once with the six hot-path attributes of `consenrich.cconsenrich`
replaced by the B200 implementations (consenrich_b200.install, the seam of tests/test_core.py:1317).

Everything between the kernels (Q0 seeding, the outer background alternation, the stopping rules)
is the reference's code in both runs, so what is compared is the public return tuple of
runConsenrich: state, covariance, residuals, NIS, block map.  The ECM stops on relative NLL
changes, so float64 re-association can in principle move an iteration count; the tolerance below
(10x the sweep tolerance) is what the public outputs are asserted to."""
import logging
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from parity_util import ATOL_REL, RTOL, assert_sweep_tracks_close

pytestmark = pytest.mark.gpu

DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver")


@pytest.fixture(scope="module")
def ref_core():
    if not os.path.isdir(os.path.join(DRIVER, "consenrich")):
        pytest.skip("oracle/_ref/driver not built (oracle/build_ref_driver.sh needs /root/reference)")
    sys.path.insert(0, DRIVER)
    import consenrich.core as core
    logging.getLogger("consenrich").setLevel(logging.ERROR)
    logging.disable(logging.WARNING)
    yield core
    logging.disable(logging.NOTSET)
    sys.path.remove(DRIVER)


def _tracks(seed, m, n):
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    x = 0.4 * np.sin(k / 700.0)
    for _ in range(max(2, n // 1500)):
        c, w, h = rng.integers(0, n), rng.uniform(8, 60), rng.uniform(0.8, 4.0)
        x = x + h * np.exp(-0.5 * ((k - c) / w) ** 2)
    v0 = rng.uniform(0.05, 0.3, size=(m, 1))
    munc = (v0 * (1 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, (m, n))).astype(np.float32)
    data = (x[None, :] + rng.normal(0, 0.05, (m, 1)) + rng.normal(size=(m, n)) * np.sqrt(munc)).astype(np.float32)
    return np.ascontiguousarray(data), np.ascontiguousarray(munc)


BASE = dict(deltaF=1.0, minQ=1e-6, maxQ=1e3, stateInit=0.0, stateCovarInit=1000.0, boundState=False,
            stateLowerBound=0.0, stateUpperBound=0.0, blockLenIntervals=100, returnDiagnostics=True)
CASES = {
    # runConsenrich's own keyword defaults (core.py:3877-3889): lambda and kappa re-weighting, background fit
    "api_defaults": dict(),
    # the CLI / YAML defaults (constants.py:266-282): lambda off, rtol 1e-6, up to 32 outer passes
    "cli_defaults": dict(ECM_fixedBackgroundRtol=1e-6, ECM_useObsPrecisionReweighting=False, ECM_outerIters=32,
                         ECM_minOuterIters=3, ECM_backgroundSmoothness=128.0, processPrecisionMultiplierMin=5e-3,
                         processPrecisionMultiplierMax=5e3),
    # deterministic budget (SURVEY 8d L-run)
    "fixed_budget": dict(ECM_fixedBackgroundRtol=0.0, ECM_fixedBackgroundIters=3, fitBackground=False, ECM_outerIters=1),
    # level-only state model
    "level_model": dict(stateModel="level", ECM_fixedBackgroundRtol=1e-5),
    # adaptive process noise (core.py:3974-3976; the reference's collected _caseRunConsenrichAPNSmoke)
    "apn": dict(ECM_useAPN=True, ECM_fixedBackgroundIters=3, ECM_outerIters=1,
                processNoiseCalibration="fixedDiagonal"),
}
# BASELINE.json configs[0] ("cfg1"): the reference's two-sample run at the default bin size.  The BAM decoding of
# that configuration is htslib work outside this path (and tests/smallTest.bam is absent from the reference
# tree, .MISSING_LARGE_BLOBS:5; pysam is not installed), so the case enters where the path does: a two-track
# count / variance matrix at the default 50 bp interval size over a smallTest-sized region, CLI defaults.
CFG1 = dict(m=2, n=40_000, kw=CASES["cli_defaults"])


@pytest.mark.parametrize("case", list(CASES) + ["cfg1_two_sample"])
def test_run_consenrich_with_b200_kernels_matches_the_reference(ref_core, case):
    import consenrich_b200 as cb
    m, n = (CFG1["m"], CFG1["n"]) if case == "cfg1_two_sample" else (6, 60_000)
    data, munc = _tracks(11 + len(case), m, n)
    kw = {**BASE, **(CFG1["kw"] if case == "cfg1_two_sample" else CASES[case])}
    want = ref_core.runConsenrich(data, munc, **kw)
    mod = cb.install()
    try:
        assert mod.cforwardPass is cb.cforwardPass  # the driver now calls into libconsenrich_b200.so
        assert mod.csolveZeroCenteredBackground is cb.csolveZeroCenteredBackground
        solves = []
        mod.csolveZeroCenteredBackground = lambda *a, **k: (solves.append(1), cb.csolveZeroCenteredBackground(*a, **k))[1]
        launches0 = cb._lib.default_context().launch_count
        got = ref_core.runConsenrich(data, munc, **kw)
        launches = cb._lib.default_context().launch_count - launches0
    finally:
        cb.uninstall()
    assert launches > 10, "runConsenrich did not reach the GPU kernels"
    if kw.get("fitBackground", True):
        assert solves, "the background update did not reach the device solve"
    state_w, P_w, res_w, nis_w, bm_w, diag_w = want
    state_g, P_g, res_g, nis_g, bm_g, diag_g = got
    np.testing.assert_array_equal(bm_g, bm_w)
    for g, w in ((state_g, state_w), (P_g, P_w), (res_g, res_w), (nis_g, nis_w)):
        assert g.shape == w.shape and g.dtype == w.dtype
    close = assert_sweep_tracks_close
    tol = dict(rtol=10 * RTOL, atol_rel=10 * ATOL_REL)
    close(state_g, state_w, f"{case} state", **tol)
    close(P_g, P_w, f"{case} stateCovar", scale="component", **tol)
    close(res_g, res_w, f"{case} residuals", **tol)
    close(nis_g, nis_w, f"{case} NIS", **tol)
    assert abs(diag_g["final_nll"] - diag_w["final_nll"]) <= 1e-5 * abs(diag_w["final_nll"])
    # report how close the two runs really are (printed with -s / on failure)
    err = np.abs(state_g.astype(np.float64) - state_w).max() / np.abs(state_w).max()
    print(f"{case}: max |state err| / scale = {err:.2e}, {launches} kernel launches, "
          f"final NLL {diag_g['final_nll']:.6f} vs {diag_w['final_nll']:.6f}")
