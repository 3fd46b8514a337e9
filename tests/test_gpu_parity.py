"""GPU parity tests: the sm_100a path, called through the C ABI (ctypes -> libconsenrich_b200.so),
against the CPU oracle on the same seeded inputs, against the committed golden fixtures produced
by the reference build, and through size-independent properties at the benchmark size.

Floating-point tolerance: tests/parity_util.py (rtol 1e-4 + 1e-5 of the track's scale); integer
outputs, Q tracks and residual identities are compared exactly.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, synth_tracks
from parity_util import assert_sweep_tracks_close, assert_tracks_close

pytestmark = pytest.mark.gpu

F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
Q0 = np.array([[2e-3, 0.0], [0.0, 1e-4]], np.float32)


@pytest.fixture(scope="module")
def cb():
    import consenrich_b200 as cb
    cb._lib.default_context(0)  # fails loudly without the library or a device
    return cb


def _weights(rng, n):
    lam = (0.1 + 5 * rng.random(n)).astype(np.float32)
    kap = np.exp(rng.normal(0, 2, n)).astype(np.float32)
    qs = (0.5 + rng.random(n)).astype(np.float32)
    qs[0] = 1.0
    return lam, kap, qs


def _sweep(mod, dim, data, munc, lam=None, kap=None, qs=None, nll_in_d=False, bounds=(0.25, 4.0, 5e-3, 5e3),
           state_init=0.25):
    m, n = data.shape
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=np.zeros(n, np.int32),
              blockCount=1, stateInit=state_init, stateCovarInit=1000.0, pad=1e-4, returnNLL=True,
              storeNLLInD=nll_in_d, lambdaExp=lam, processPrecExp=kap, processQScale=qs,
              obsPrecisionMultiplierMin=bounds[0], obsPrecisionMultiplierMax=bounds[1],
              procPrecisionMultiplierMin=bounds[2], procPrecisionMultiplierMax=bounds[3])
    o = dict(xf=np.empty((n, dim), np.float32), Pf=np.empty((n, dim, dim), np.float32),
             Qf=np.zeros((n, dim, dim), np.float32), D=np.empty(n, np.float32))
    st = dict(stateForward=o["xf"], stateCovarForward=o["Pf"], pNoiseForward=o["Qf"], vectorD=o["D"])
    if dim == 2:
        r = mod.cforwardPass(matrixF=F, **kw, **st)
        b = mod.cbackwardPass(matrixData=data, matrixF=F, stateForward=o["xf"], stateCovarForward=o["Pf"],
                              pNoiseForward=o["Qf"])
    else:
        r = mod.cforwardPassLevel(**kw, **st)
        b = mod.cbackwardPassLevel(matrixData=data, stateForward=o["xf"], stateCovarForward=o["Pf"],
                                   pNoiseForward=o["Qf"])
    assert r[2] is o["D"]  # the supplied vectorD object is returned (tests/test_core.py:3427)
    o.update(phi=r[0], nll=r[3], xs=b[0], Ps=b[1], lag=b[2], res=b[3])
    return o


def _compare_sweeps(got, want, n, dim, label):
    close = assert_sweep_tracks_close
    close(got["xf"], want["xf"], f"{label} stateForward")
    close(got["Pf"], want["Pf"], f"{label} stateCovarForward", scale="component")
    np.testing.assert_array_equal(got["Qf"][: n - 1], want["Qf"][: n - 1], err_msg=f"{label} pNoiseForward")
    close(got["D"], want["D"], f"{label} vectorD")
    assert abs(got["nll"] - want["nll"]) <= 2e-6 * max(abs(want["nll"]), 1.0), (label, got["nll"], want["nll"])
    assert abs(got["phi"] - want["phi"]) <= 1e-4 * max(abs(want["phi"]), 1e-3), label
    close(got["xs"], want["xs"], f"{label} stateSmoothed")
    close(got["Ps"], want["Ps"], f"{label} stateCovarSmoothed", scale="component")
    if n > 1:
        close(got["lag"], want["lag"], f"{label} lagCovSmoothed", scale="component")
    close(got["res"], want["res"], f"{label} postFitResiduals")
    if dim == 1:  # the level filter carries float64 in the reference as well
        assert_tracks_close(got["xf"], want["xf"], f"{label} level stateForward", rtol=1e-6, atol_rel=1e-7)


SWEEP_CASES = [
    # m, n, masked_frac, weights
    (2, 1, 0.0, False), (2, 2, 0.0, True), (3, 3, 0.0, False), (4, 7, 0.0, True), (3, 257, 0.0, False),
    (10, 1023, 0.05, True), (10, 1024, 0.0, False), (7, 1025, 0.0, True), (5, 2048, 0.3, True),
    (33, 4097, 0.0, True), (1, 5000, 0.0, False), (130, 3000, 0.1, True),
    (10, 300001, 0.02, True), (3, 1200003, 0.0, False),
]


@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("case", SWEEP_CASES)
def test_sweep_matches_oracle(cb, oracle, dim, case):
    m, n, masked, weights = case
    data, munc = synth_tracks(5000 + n + m, m, n, masked_frac=masked)
    lam = kap = qs = None
    if weights:
        lam, kap, qs = _weights(np.random.default_rng(n), n)
    want = _sweep(oracle, dim, data, munc, lam, kap, qs)
    got = _sweep(cb, dim, data, munc, lam, kap, qs)
    _compare_sweeps(got, want, n, dim, f"{m}x{n} d{dim}")
    # exact identities of the residual track
    lvl = got["xs"][:, 0].astype(np.float64)
    np.testing.assert_array_equal(got["res"], (data.T.astype(np.float64) - lvl[:, None]).astype(np.float32))
    np.testing.assert_array_equal(got["xs"][-1], got["xf"][-1])
    np.testing.assert_array_equal(got["Ps"][-1], got["Pf"][-1])


@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("nsub", [1, 2, 3, 5, 8, 16])
def test_sweep_is_independent_of_the_run_length(cb, oracle, dim, nsub):
    """Every thread of the scan kernels runs through nsub sub-steps of 4 bins; the launch heuristic
    picks nsub from the track length.  Force each value: ragged tails, runs that straddle the end,
    several tiles per value."""
    from consenrich_b200 import _lib
    L = _lib.load()
    try:
        _lib.check(L.cb200_set_scan_substeps(nsub))
        sizes = [(4, 4 * 128 * nsub * 3 + 13), (6, 4 * 128 * nsub - 1), (3, 4 * nsub + 1), (9, 70001)]
        if nsub == 1:
            sizes.append((2, 1_300_003))  # 2540 tiles: the look-back goes through several level-1 rounds
        for m, n in sizes:
            data, munc = synth_tracks(900 + n + nsub, m, n, masked_frac=0.03)
            lam, kap, qs = _weights(np.random.default_rng(n), n)
            want = _sweep(oracle, dim, data, munc, lam, kap, qs)
            got = _sweep(cb, dim, data, munc, lam, kap, qs)
            _compare_sweeps(got, want, n, dim, f"nsub{nsub} {m}x{n} d{dim}")
    finally:
        _lib.check(L.cb200_set_scan_substeps(0))


@pytest.mark.parametrize("dim", [2, 1])
def test_smoother_alone_on_identical_forward_tracks(cb, oracle, dim):
    """cbackwardPass fed the ORACLE's float32 forward tracks: isolates the reverse scan, so even the
    steady rows must agree ten times tighter than the stated tolerance."""
    for m, n in ((130, 3000), (5, 20000), (3, 1)):
        data, munc = synth_tracks(42 + n, m, n, masked_frac=0.1)
        lam, kap, qs = _weights(np.random.default_rng(n), n)
        want = _sweep(oracle, dim, data, munc, lam, kap, qs)
        kw = dict(matrixData=data, stateForward=want["xf"], stateCovarForward=want["Pf"], pNoiseForward=want["Qf"])
        b = cb.cbackwardPass(matrixF=F, **kw) if dim == 2 else cb.cbackwardPassLevel(**kw)
        # transient rows: the reference's un-pivoted 2x2 inverse of P^- ~ 500 [[1,1],[1,1]] loses ~1e-4
        # there even in float64 (FMA contraction alone moves it), hence the transient factor
        close = assert_sweep_tracks_close
        close(b[0], want["xs"], "stateSmoothed", rtol=1e-5, atol_rel=1e-6)
        close(b[1], want["Ps"], "stateCovarSmoothed", scale="component", rtol=1e-5, atol_rel=1e-6)
        if n > 1:
            close(b[2], want["lag"], "lagCovSmoothed", scale="component", rtol=1e-5, atol_rel=1e-6)
        close(b[3], want["res"], "postFitResiduals", rtol=1e-5, atol_rel=1e-6)


def test_sweep_nll_in_d_and_cli_bounds(cb, oracle):
    m, n = 6, 5000
    data, munc = synth_tracks(77, m, n)
    lam, kap, qs = _weights(np.random.default_rng(3), n)
    for dim in (2, 1):
        want = _sweep(oracle, dim, data, munc, lam, kap, qs, nll_in_d=True, bounds=(0.5, 2.0, 5e-3, 5e3))
        got = _sweep(cb, dim, data, munc, lam, kap, qs, nll_in_d=True, bounds=(0.5, 2.0, 5e-3, 5e3))
        _compare_sweeps(got, want, n, dim, f"nll_in_d d{dim}")


def _golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
    cases = {}
    for key in z.files:
        case, rest = key.split("/", 1)
        cases.setdefault(case, {})[rest] = z[key]
    return cases


@pytest.mark.parametrize("dim", [2, 1])
def test_sweep_matches_reference_golden_vectors(cb, dim):
    """Fixtures generated by the unmodified reference build (tests/golden/make_golden.py)."""
    for name, case in _golden("sweep_golden.npz").items():
        data, munc, Q, bm = case["data"], case["munc"], case["Q0"], case["blockMap"]
        n = data.shape[1]
        extra = {k[6:]: (v if v.ndim else v.item()) for k, v in case.items() if k.startswith("extra/")}
        kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q, intervalToBlockMap=bm,
                  blockCount=int(bm.max()) + 1, stateInit=float(case["stateInit"]), stateCovarInit=1000.0,
                  pad=1.0e-4, returnNLL=True, **extra)
        st = dict(stateForward=np.empty((n, dim), np.float32), stateCovarForward=np.empty((n, dim, dim), np.float32),
                  pNoiseForward=np.zeros((n, dim, dim), np.float32), vectorD=np.empty(n, np.float32))
        apn = bool(extra.get("ECM_useAPN"))
        if dim == 2:
            r = cb.cforwardPass(matrixF=F, **kw, **st)
            b = cb.cbackwardPass(matrixData=data, matrixF=F, stateForward=st["stateForward"],
                                 stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"])
        else:
            r = cb.cforwardPassLevel(**kw, **st)
            b = cb.cbackwardPassLevel(matrixData=data, stateForward=st["stateForward"],
                                      stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"])
        pre = f"d{dim}/"
        close = assert_sweep_tracks_close
        close(st["stateForward"], case[pre + "stateForward"], name)
        close(st["stateCovarForward"], case[pre + "stateCovarForward"], name, scale="component")
        if apn:  # the process noise follows the innovation statistic through the APN feedback (pyx:510-527)
            close(st["pNoiseForward"][: n - 1], case[pre + "pNoiseForward"][: n - 1], name, scale="component")
        else:
            np.testing.assert_array_equal(st["pNoiseForward"][: n - 1], case[pre + "pNoiseForward"][: n - 1])
        close(st["vectorD"], case[pre + "vectorD"], name)
        assert abs(r[3] - float(case[pre + "sumNLL"])) <= 2e-6 * max(abs(float(case[pre + "sumNLL"])), 1.0)
        close(b[0], case[pre + "stateSmoothed"], name)
        close(b[1], case[pre + "stateCovarSmoothed"], name, scale="component")
        if n > 1:
            close(b[2], case[pre + "lagCovSmoothed"], name, scale="component")
        close(b[3], case[pre + "postFitResiduals"], name)


def _ecm(mod, dim, data, munc, **opts):
    n = data.shape[1]
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=np.zeros(n, np.int32),
              blockCount=1, stateInit=0.0, stateCovarInit=1000.0, returnIntermediates=True, returnDiagnostics=True,
              logIterations=False, **opts)
    return mod.cfixedBackgroundECM(matrixF=F, **kw) if dim == 2 else mod.cfixedBackgroundECMLevel(**kw)


# Stated ECM tolerance, set from the measured GPU-vs-oracle error over every case of this file and of
# test_lean_sweeps.py (tools/ecm_error_probe.py on B200: state <= 3.2e-7, residuals <= 4.2e-7, covariances
# <= 1.8e-5, lambda <= 7.7e-7, kappa <= 1.9e-6 of the track's scale; final NLL <= 3.8e-9 relative): about ten
# times those.  The ECM compounds the sweeps' re-association error through the multiplier updates, so it
# sits above the single-sweep figure (tests/parity_util.py) but far below the 1e-4 the stopping rule could
# introduce if iteration counts diverged -- and those are compared exactly.
ECM_TOL = dict(state=dict(rtol=2e-5, atol_rel=4e-6), cov=dict(rtol=2e-4, atol_rel=1e-4),
               mult=dict(rtol=5e-5, atol_rel=1e-5), nll=5e-8)
# Where the two sides may legitimately take different discrete decisions -- a free-running stopping rule that
# ends on different iterations, or the adaptive-process-noise feedback, which branches on D_k against a
# threshold (pyx:510-527) -- the comparison is at the level such a decision moves the result.
ECM_TOL_DISCRETE = dict(state=dict(rtol=2e-4, atol_rel=1e-4), cov=dict(rtol=2e-3, atol_rel=1e-4),
                        mult=dict(rtol=2e-3, atol_rel=1e-4), nll=1e-6)


def _compare_ecm(a, b, label, exact_iters=True, tol=None):
    if exact_iters:
        assert a[0] == b[0], label
    if tol is None:
        tol = ECM_TOL if a[0] == b[0] else ECM_TOL_DISCRETE
    assert abs(a[1] - b[1]) <= tol["nll"] * max(abs(b[1]), 1.0), f"{label}: nll {a[1]} vs {b[1]}"
    assert_tracks_close(a[2], b[2], f"{label} stateSmoothed", **tol["state"])
    assert_tracks_close(a[3], b[3], f"{label} stateCovarSmoothed", scale="component", **tol["cov"])
    assert_tracks_close(a[4], b[4], f"{label} lagCovSmoothed", scale="component", **tol["cov"])
    assert_tracks_close(a[5], b[5], f"{label} residuals", **tol["state"])
    for x, y, nm in ((a[6], b[6], "lambda"), (a[7], b[7], "kappa")):
        assert (x is None) == (y is None), f"{label} {nm}"
        if x is not None:
            assert_tracks_close(x, y, f"{label} {nm}", **tol["mult"])
    da, db = a[8], b[8]
    assert set(da) == set(db)
    for key in ("max_iters", "skipped", "skip_reason", "fallback", "patience_target"):
        assert da[key] == db[key], (label, key)


@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("opts", [
    dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2),
    dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, t_innerIters=3,
         ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3),
    dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, ECM_useProcessPrecisionReweighting=False),
    dict(ECM_fixedBackgroundIters=4, ECM_fixedBackgroundRtol=0.0, t_innerIters=1, ECM_robustTNu=4.0,
         trackOptimizationPath=True),
])
def test_ecm_fixed_budget_matches_oracle(cb, oracle, dim, opts):
    """Deterministic iteration budget (rtol = 0): the multipliers go through an identical number of
    updates on both sides, so the tracks are comparable without the stopping rule's discreteness."""
    data, munc = synth_tracks(909, 8, 6000, masked_frac=0.02)
    a, b = _ecm(cb, dim, data, munc, **opts), _ecm(oracle, dim, data, munc, **opts)
    _compare_ecm(a, b, f"ecm d{dim} {sorted(opts)}")
    if opts.get("trackOptimizationPath"):
        pa, pb = a[8]["optimization_path"], b[8]["optimization_path"]
        assert len(pa) == len(pb)
        for ea, eb in zip(pa, pb):
            assert set(ea) == set(eb) and ea["iter"] == eb["iter"] and ea["reset_iteration"] == eb["reset_iteration"]


@pytest.mark.parametrize("dim", [2, 1])
def test_ecm_free_running_and_qscale_and_warm_start(cb, oracle, dim):
    data, munc = synth_tracks(31, 5, 3000)
    rng = np.random.default_rng(5)
    qs = (0.5 + rng.random(3000)).astype(np.float32)
    qs[0] = 1.0
    opts = dict(ECM_fixedBackgroundIters=25, ECM_fixedBackgroundRtol=1e-4, processQScale=qs,
                lambdaExpInit=(0.2 + 6 * rng.random(3000)).astype(np.float32),
                processPrecExpInit=np.exp(rng.normal(0, 1, 3000)).astype(np.float32))
    a, b = _ecm(cb, dim, data, munc, **opts), _ecm(oracle, dim, data, munc, **opts)
    _compare_ecm(a, b, f"ecm free d{dim}")
    assert a[8]["converged"] == b[8]["converged"]


@pytest.mark.parametrize("dim", [2, 1])
def test_ecm_tiny_track_uses_filter_smoother_fallback(cb, oracle, dim):
    """n <= 5 (cconsenrich.pyx:7998-8129; reference case _caseCFixedBackgroundECMTinyTrackUsesFiniteFallback)."""
    data, munc = synth_tracks(3, 3, 4)
    a, b = _ecm(cb, dim, data, munc), _ecm(oracle, dim, data, munc)
    assert a[0] == b[0] == 0
    _compare_ecm(a, b, f"tiny d{dim}")
    assert a[8]["skipped"] and a[8]["fallback"] == "filter_smoother_only"
    np.testing.assert_array_equal(a[6], b[6])
    np.testing.assert_array_equal(a[7], b[7])
    short = (cb.cfixedBackgroundECM(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=Q0,
                                    intervalToBlockMap=np.zeros(4, np.int32), blockCount=1, stateInit=0.0,
                                    stateCovarInit=1000.0))
    assert len(short) == 2 and short[0] == 0


def test_disabled_multipliers_are_returned_as_none(cb):
    data, munc = synth_tracks(8, 3, 200)
    out = _ecm(cb, 2, data, munc, ECM_fixedBackgroundIters=1, ECM_useObsPrecisionReweighting=False,
               ECM_useProcessPrecisionReweighting=False)
    assert out[6] is None and out[7] is None  # tests/test_core.py:3078-3079


def test_error_behaviour_matches_reference(cb, oracle):
    data, munc = synth_tracks(1, 3, 50)
    base = dict(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=Q0, blockCount=1, stateInit=0.0,
                stateCovarInit=1000.0)
    bad_bm = np.zeros(50, np.int32)
    bad_bm[17] = 3
    bad_qs = np.ones(50, np.float32)
    bad_qs[0] = 2.0
    for extra in (dict(intervalToBlockMap=bad_bm), dict(intervalToBlockMap=np.zeros(50, np.int32), blockCount=0),
                  dict(intervalToBlockMap=np.zeros(50, np.int32), processQScale=bad_qs),
                  dict(intervalToBlockMap=np.zeros(50, np.int32), obsPrecisionMultiplierMin=0.0),
                  dict(intervalToBlockMap=np.zeros(10, np.int32)),
                  dict(intervalToBlockMap=np.zeros(50, np.int32), matrixPluginMuncInit=munc[:, :40].copy())):
        kw = {**base, **extra}
        with pytest.raises(ValueError) as want:
            oracle.cforwardPass(**kw)
        with pytest.raises(ValueError) as got:
            cb.cforwardPass(**kw)
        assert str(got.value) == str(want.value)
    with pytest.raises(ValueError, match="matrixQ0 is singular"):
        cb.cfixedBackgroundECM(**{**base, "intervalToBlockMap": np.zeros(50, np.int32),
                                  "matrixQ0": np.ones((2, 2), np.float32)})
    # empty input returns zeros without touching the device (pyx:6494-6501)
    e = np.empty((3, 0), np.float32)
    r = cb.cforwardPass(**{**base, "matrixData": e, "matrixPluginMuncInit": e}, intervalToBlockMap=np.zeros(0, np.int32),
                        returnNLL=True)
    assert r[0] == 0.0 and r[1] == 0 and r[2].shape == (0,) and r[3] == 0.0


def test_install_routes_the_reference_seam(cb):
    """install() swaps the six attributes on a module object, as the reference's tests do with
    monkeypatch.setattr(cconsenrich, ...) (tests/test_core.py:1300-1317)."""
    import types
    fake = types.ModuleType("cconsenrich")
    fake.cforwardPass = lambda *a, **k: "reference"
    cb.install(fake)
    assert fake.cforwardPass is cb.cforwardPass and fake.cfixedBackgroundECMLevel is cb.cfixedBackgroundECMLevel
    cb.uninstall(fake)
    assert fake.cforwardPass() == "reference"


# ---------------------------------------------------------------------------------------------
# benchmark-size checks (BASELINE.json configs[1]: m = 10, hg38 chr19 at 25 bp = 2 344 705 bins)
# ---------------------------------------------------------------------------------------------
CHR19_BINS_25BP = 2344705


def test_chr19_sized_sweep_matches_oracle_and_properties(cb, oracle):
    m, n = 10, CHR19_BINS_25BP
    rng = np.random.default_rng(1729)
    k = np.arange(n, dtype=np.float64)
    x = 0.5 * np.sin(2 * np.pi * k / 5.0e4)
    centers = rng.integers(0, n, size=3000)
    for c, w, h in zip(centers, rng.uniform(8, 80, 3000), rng.uniform(0.5, 4.0, 3000)):
        lo, hi = max(0, int(c - 5 * w)), min(n, int(c + 5 * w))
        x[lo:hi] += h * np.exp(-0.5 * ((k[lo:hi] - c) / w) ** 2)
    v0 = rng.uniform(0.05, 0.3, size=(m, 1))
    munc = (v0 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n))).astype(np.float32)
    data = (x[None, :] + rng.normal(0, 0.05, size=(m, 1)) + rng.normal(size=(m, n)) * np.sqrt(munc)).astype(np.float32)
    want = _sweep(oracle, 2, data, munc)
    got = _sweep(cb, 2, data, munc)
    _compare_sweeps(got, want, n, 2, "chr19")
    # size-independent properties
    lvl = got["xs"][:, 0].astype(np.float64)
    np.testing.assert_array_equal(got["res"], (data.T.astype(np.float64) - lvl[:, None]).astype(np.float32))
    assert np.all(got["Ps"][:, 0, 0] <= got["Pf"][:, 0, 0] * (1 + 1e-5) + 1e-12)  # smoothing never adds variance
    assert np.all(got["Pf"][:, 0, 0] > 0) and np.all(got["Ps"][:, 1, 1] >= 0)
    np.testing.assert_array_equal(got["Ps"][:, 0, 1], got["Ps"][:, 1, 0])
    again = _sweep(cb, 2, data, munc)  # look-back windows differ run to run; results may not
    assert_tracks_close(again["xs"], got["xs"], "rerun", rtol=1e-6, atol_rel=1e-6)
    # the filter is linear in (data, stateInit): scaling both by 2 scales the states by 2 exactly
    # in exact arithmetic (powers of two commute with rounding)
    twice = _sweep(cb, 2, 2.0 * data, munc, state_init=0.5)
    assert_tracks_close(twice["xs"], 2.0 * got["xs"], "linearity", rtol=1e-5, atol_rel=1e-6)
    assert_tracks_close(twice["Ps"], got["Ps"], "covariance is data-independent", scale="component",
                        rtol=1e-6, atol_rel=1e-7)


def test_longest_chromosome_at_10bp_and_widest_cohort(cb, oracle):
    """Maximum sizes of BASELINE.json's configs: chr1 at 10 bp (24 895 643 bins, cfg4's longest track:
    ~3040 tiles, several waves, multi-round look-back, 64-bit offsets) with few tracks, and a
    1000-track cohort (cfg5's width: 32 track tiles in the residual transpose, long fold loops with
    mantissa renormalisation)."""
    # ---- longest chromosome ----
    m, n = 2, 24_895_643
    rng = np.random.default_rng(5)
    k = np.arange(n, dtype=np.float32)
    x = (0.5 * np.sin(k / 9000.0) + 1.5 * (np.sin(k / 411.0) > 0.97)).astype(np.float32)
    munc = (0.15 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n)).astype(np.float32)).astype(np.float32)
    data = (x[None, :] + rng.standard_normal((m, n), dtype=np.float32) * np.sqrt(munc)).astype(np.float32)
    want = _sweep(oracle, 2, data, munc)
    got = _sweep(cb, 2, data, munc)
    _compare_sweeps(got, want, n, 2, "chr1@10bp")
    np.testing.assert_array_equal(got["xs"][-1], got["xf"][-1])
    del want, got
    # ---- widest cohort ----
    m, n = 1000, 6_001
    data, munc = synth_tracks(31337, m, n, masked_frac=0.05)
    lam, kap, qs = _weights(np.random.default_rng(3), n)
    for dim in (2, 1):
        want = _sweep(oracle, dim, data, munc, lam, kap, qs)
        got = _sweep(cb, dim, data, munc, lam, kap, qs)
        _compare_sweeps(got, want, n, dim, f"1000 tracks d{dim}")
        lvl = got["xs"][:, 0].astype(np.float64)
        np.testing.assert_array_equal(got["res"], (data.T.astype(np.float64) - lvl[:, None]).astype(np.float32))


@pytest.mark.parametrize("nsub", [4, 7, 16])
@pytest.mark.parametrize("opts", [
    # the CLI configuration: kappa only -> lean inner sweeps, fused kappa update, run-ahead NLL pass
    dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2,
         ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3),
    # lambda and kappa, free-running with a tolerance (stops early: the run-ahead pass is dropped)
    dict(ECM_fixedBackgroundIters=12, ECM_fixedBackgroundRtol=1e-3, t_innerIters=2),
])
def test_ecm_with_run_elements_composed_by_the_forward_replay(cb, oracle, nsub, opts):
    """With runs of at least 16 bins the 2-state ECM lets the forward replay compose the smoother's run
    elements and the backward scan skips its first pass.  The launch heuristic only picks such runs for
    tracks of ~5e5 bins and more; force them on a track the oracle finishes quickly (several tiles, a
    ragged last run, masked cells) and on one long enough for the heuristic itself."""
    from consenrich_b200 import _lib
    L = _lib.load()
    try:
        _lib.check(L.cb200_set_scan_substeps(nsub))
        for m, n in ((7, 40_003), (3, 4 * nsub * 128 * 2 + 1)):
            data, munc = synth_tracks(4000 + n, m, n, masked_frac=0.03)
            a, b = _ecm(cb, 2, data, munc, **opts), _ecm(oracle, 2, data, munc, **opts)
            _compare_ecm(a, b, f"fused nsub{nsub} {m}x{n}", exact_iters="ECM_fixedBackgroundRtol" in opts and opts["ECM_fixedBackgroundRtol"] == 0.0)
    finally:
        _lib.check(L.cb200_set_scan_substeps(0))
    if nsub == 16:  # the heuristic's own choice on a long track
        data, munc = synth_tracks(77, 2, 900_001)
        o = dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, t_innerIters=2,
                 ECM_useObsPrecisionReweighting=False)
        _compare_ecm(_ecm(cb, 2, data, munc, **o), _ecm(oracle, 2, data, munc, **o), "fused long track")


# ---------------------------------------------------------------------------------------------
# adaptive process noise (cconsenrich.pyx:510-527, 688-703): the forward pass as a sequential device
# recursion (csrc/apn_kernels.cu)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("m,n", [(3, 1), (4, 33), (6, 1000), (5, 20_011)])
def test_adaptive_process_noise_forward_matches_oracle(cb, oracle, dim, m, n):
    data, munc = synth_tracks(600 + n, m, n, masked_frac=0.02 if n > 100 else 0.0)
    data[:, n // 2: n // 2 + 7] += 6.0  # a stretch the model does not explain: D above the threshold, Q scaled up
    lam = (0.3 + 3 * np.random.default_rng(n).random(n)).astype(np.float32)
    for extra in (dict(), dict(lambdaExp=lam, APN_dStatThresh=1.5, APN_dStatScale=4.0, APN_dStatPC=1.5, APN_minQ=1e-5,
                               APN_maxQ=10.0, storeNLLInD=False)):
        kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=np.zeros(n, np.int32),
                  blockCount=1, stateInit=0.25, stateCovarInit=1000.0, pad=1e-4, returnNLL=True, ECM_useAPN=True, **extra)
        outs = []
        for mod in (cb, oracle):
            st = dict(stateForward=np.empty((n, dim), np.float32), stateCovarForward=np.empty((n, dim, dim), np.float32),
                      pNoiseForward=np.zeros((n, dim, dim), np.float32), vectorD=np.empty(n, np.float32))
            r = mod.cforwardPass(matrixF=F, **kw, **st) if dim == 2 else mod.cforwardPassLevel(**kw, **st)
            outs.append((r, st))
        (rg, sg), (ro, so) = outs
        label = f"apn d{dim} {m}x{n} {sorted(extra)}"
        assert_sweep_tracks_close(sg["stateForward"], so["stateForward"], label + " stateForward")
        assert_sweep_tracks_close(sg["stateCovarForward"], so["stateCovarForward"], label + " stateCovarForward",
                                  scale="component")
        if n > 1:
            assert_sweep_tracks_close(sg["pNoiseForward"][: n - 1], so["pNoiseForward"][: n - 1], label + " pNoiseForward",
                                      scale="component")
            assert len(np.unique(so["pNoiseForward"][: n - 1, 0, 0])) > 1 or n < 30  # the feedback did act
        assert_sweep_tracks_close(sg["vectorD"], so["vectorD"], label + " vectorD")
        assert abs(rg[3] - ro[3]) <= 2e-6 * max(abs(ro[3]), 1.0), (label, rg[3], ro[3])
        assert abs(rg[0] - ro[0]) <= 1e-4 * max(abs(ro[0]), 1e-3), label


@pytest.mark.parametrize("dim", [2, 1])
def test_adaptive_process_noise_ecm_matches_oracle(cb, oracle, dim):
    """cfixedBackgroundECM with ECM_useAPN: kappa is not fitted (pyx:7268, 8244), lambda is; the reference's
    collected case _caseRunConsenrichAPNSmoke (tests/test_core.py:6051) goes through this path."""
    data, munc = synth_tracks(77, 6, 5_003, masked_frac=0.02)
    opts = dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2, ECM_useAPN=True)
    a, b = _ecm(cb, dim, data, munc, **opts), _ecm(oracle, dim, data, munc, **opts)
    _compare_ecm(a, b, f"ecm apn d{dim}", tol=ECM_TOL_DISCRETE)
    assert a[7] is None and b[7] is None and a[6] is not None
