"""Pins the CPU restatement of the observation-noise stage's dense kernels (oracle/munc_oracle.c):
bit-exact against golden vectors of the reference build (tests/golden/make_munc_golden.py, the
reference's own known-answer case included) and against the reference build itself when present."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR


def golden_cases(kind="smooth"):
    z = np.load(os.path.join(GOLDEN_DIR, "munc_golden.npz"), allow_pickle=False)
    cases = {}
    for key in z.files:
        k, name, field = key.split("/")
        if k == kind:
            cases.setdefault(name, {})[field] = z[key]
    return cases


def run_finalize(mod, c):
    opts = {k[4:]: (v.item() if v.ndim == 0 else v) for k, v in c.items() if k.startswith("opt_")}
    return mod.cFinalizeMuncEBTrack(c["local"], priorVarianceTrack=c.get("prior"), countFloor=c.get("countFloor"), **opts)


def check_finalize(got, c, exact=True):
    out, diag = got
    np.testing.assert_array_equal(out, c["out"])
    for k, v in c.items():
        if k.startswith("diag_"):
            assert diag[k[5:]] == v.item(), k


def run_case(mod, c):
    return mod.cMuncSmoothDenseLocalEvidence(c["local"], int(c["window"]), excludeMask=c.get("mask"),
                                             eps=float(c["eps"]))


def test_oracle_matches_golden_munc_vectors_bitwise(oracle):
    cases = golden_cases()
    assert len(cases) >= 7
    for name, c in cases.items():
        np.testing.assert_array_equal(run_case(oracle, c), c["out"], err_msg=name)
    # the reference's own expectation for its known-answer case (tests/test_core.py:1548-1556)
    want = np.asarray([[2.0, 2.0, 4.0, 5.0, 5.0], [5.0, 5.0, 12.0, 56.0 / 3.0, 56.0 / 3.0]], np.float32)
    np.testing.assert_allclose(run_case(oracle, cases["ref_case"]), want, rtol=1e-6, atol=1e-6)


def test_oracle_matches_reference_build_on_fresh_seeds(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    from golden.make_munc_golden import exclude_mask, local_evidence
    rng = np.random.default_rng(8)
    for m, n, w, mode in ((1, 1, 3, 0), (3, 2, 1, 1), (2, 33, 5, 2), (5, 2000, 17, 1), (3, 10001, 300, 2), (2, 500, 5000, 0)):
        le, mk = local_evidence(rng, m, n), exclude_mask(rng, m, n, mode)
        a = ref.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=1e-6)
        b = oracle.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=1e-6)
        np.testing.assert_array_equal(a, b)
    le = local_evidence(rng, 2, 40)
    le[1, 7] = 0.0
    for mod in (ref, oracle):
        with pytest.raises(ValueError, match="active local evidence cells must be positive and finite"):
            mod.cMuncSmoothDenseLocalEvidence(le, 5)
        mk = np.zeros(40, np.uint8)
        mk[7] = 1  # the offending cell is masked: accepted
        assert mod.cMuncSmoothDenseLocalEvidence(le, 5, excludeMask=mk).shape == (2, 40)
        with pytest.raises(ValueError, match="windowIntervals must be positive"):
            mod.cMuncSmoothDenseLocalEvidence(le, 0)
        with pytest.raises(ValueError, match="eps must be positive and finite"):
            mod.cMuncSmoothDenseLocalEvidence(le, 3, eps=0.0)
        with pytest.raises(ValueError, match="excludeMask length must match interval count"):
            mod.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=np.zeros(39, np.uint8))
        with pytest.raises(ValueError, match="excludeMask shape must match localEvidence shape"):
            mod.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=np.zeros((3, 40), np.uint8))


def test_oracle_matches_golden_finalize_vectors_bitwise(oracle):
    cases = golden_cases("finalize")
    assert len(cases) >= 5
    for name, c in cases.items():
        check_finalize(run_finalize(oracle, c), c)
    out, diag = run_finalize(oracle, cases["ref_case"])  # the reference's own expectation
    np.testing.assert_allclose(out, np.asarray([0.23, 0.69, 5.0, 0.016], np.float32), rtol=1e-6, atol=1e-6)
    assert (diag["supportCount"], diag["countFloorFiniteCount"], diag["countFloorAddedCount"],
            diag["countFloorMissingCount"], diag["finalShrinkagePairCount"]) == (3, 3, 2, 1, 4)


def test_oracle_finalize_matches_reference_build_errors_included(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(12)
    n = 4000
    loc = rng.uniform(1e-4, 3.0, n).astype(np.float32)
    pri = rng.uniform(1e-4, 3.0, n).astype(np.float32)
    cf = rng.uniform(0, 1, n).astype(np.float32)
    cf[::7] = np.nan
    kw = dict(nuLocal=11.0, nuPrior=5.5, varianceFloor=1e-3, varianceCap=2.5)
    a, b = ref.cFinalizeMuncEBTrack(loc, pri, cf, **kw), oracle.cFinalizeMuncEBTrack(loc, pri, cf, **kw)
    np.testing.assert_array_equal(a[0], b[0])
    assert a[1] == b[1]
    bad_l, bad_p, bad_c = loc.copy(), pri.copy(), cf.copy()
    bad_l[900] = 0.0
    bad_p[300] = np.inf
    bad_c[300] = -1.0
    for args in ((bad_l, pri, cf), (loc, bad_p, cf), (loc, pri, bad_c), (bad_l, bad_p, bad_c), (loc, bad_p, bad_c)):
        msgs = []
        for mod in (ref, oracle):
            with pytest.raises(ValueError) as e:
                mod.cFinalizeMuncEBTrack(*args, **kw)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1], msgs
    for bad_kw in (dict(varianceFloor=0.0), dict(varianceCap=1e-4), dict(nuLocal=0.0), dict(nuPrior=float("nan"))):
        msgs = []
        for mod in (ref, oracle):
            with pytest.raises(ValueError) as e:
                mod.cFinalizeMuncEBTrack(loc, pri, cf, **{**kw, **bad_kw})
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1]


def run_seed(mod, c):
    ins = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in c.items() if k.startswith("in_")}
    pos = [ins.pop(k) for k in ("matrixData", "matrixMunc", "stateMean", "stateVariance")]
    return mod.cMuncObservationMomentSeedPass(*pos, **ins)


def check_seed(got, c, what=""):
    for k, v in zip(("moment", "rhoOut", "omegaRaw", "omegaOut", "local", "variance"), got):
        np.testing.assert_array_equal(v, c["out_" + k], err_msg=f"{what} {k}")


def test_oracle_matches_golden_seed_pass_vectors_bitwise(oracle):
    cases = golden_cases("seed")
    assert len(cases) >= 5
    for name, c in cases.items():
        check_seed(run_seed(oracle, c), c, name)


def test_oracle_seed_pass_matches_reference_build_errors_included(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    from golden.make_munc_golden import SEED_POSITIONAL, seed_case
    rng = np.random.default_rng(21)
    for variant in ("update", "fixed", "unweighted", "gaussian"):
        for m, n in ((1, 3), (7, 2500)):
            c = seed_case(rng, m, n, variant)
            pos = [c[k] for k in SEED_POSITIONAL]
            kw = {k: v for k, v in c.items() if k not in SEED_POSITIONAL}
            for x, y in zip(ref.cMuncObservationMomentSeedPass(*pos, **kw), oracle.cMuncObservationMomentSeedPass(*pos, **kw)):
                np.testing.assert_array_equal(x, y)
    c = seed_case(rng, 3, 50, "update")
    pos = [c[k] for k in SEED_POSITIONAL]
    kw = {k: v for k, v in c.items() if k not in SEED_POSITIONAL}
    k_act = int(np.flatnonzero(c["activeMask"])[0])
    bad_data = pos[0].copy()
    bad_data[1, k_act] = np.nan
    bad_munc = pos[1].copy()
    bad_munc[2, k_act] = -1.0
    for args, kwargs in (((bad_data, *pos[1:]), kw), ((pos[0], bad_munc, *pos[2:]), kw), (pos, {**kw, "pad": -1.0}),
                         (pos, {**kw, "varianceFloor": 0.0}), (pos, {**kw, "omegaMin": 0.0}),
                         (pos, {**kw, "omegaIn": c["omegaIn"][:-1]}), (pos, {**kw, "countFloor": c["countFloor"][:, :-1]})):
        msgs = []
        for mod in (ref, oracle):
            with pytest.raises(ValueError) as e:
                mod.cMuncObservationMomentSeedPass(*args, **kwargs)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1], msgs


def test_oracle_matches_golden_ema_vectors_bitwise(oracle):
    cases = golden_cases("ema")
    assert len(cases) >= 7
    for name, c in cases.items():
        got = oracle.cEMA(c["x"], float(c["alpha"]))
        assert got.dtype == c["out"].dtype
        np.testing.assert_array_equal(got, c["out"], err_msg=name)


def test_oracle_ema_matches_reference_build(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(2)
    for dt in (np.float32, np.float64):
        for n, alpha in ((1, 0.3), (2, 0.9), (257, 0.05), (50_001, 2.0 / 300.0), (4000, 0.0), (4000, 1.0)):
            x = rng.normal(size=n).astype(dt)
            np.testing.assert_array_equal(ref.cEMA(x, alpha), oracle.cEMA(x, alpha))
    assert oracle.cEMA(np.arange(5), 0.5).dtype == np.float64  # everything but float32 is filtered as float64


def test_installed_smoother_hands_oversized_windows_to_the_function_it_replaced():
    """No GPU needed: a window beyond the device kernel's limit (a configuration the reference accepts) goes to
    the replaced host function under install(), with a HostPathWarning -- never a failure, never silent."""
    import types

    import consenrich_b200 as cb
    from consenrich_b200 import native
    calls = []
    fake = types.ModuleType("cconsenrich")
    fake.cMuncSmoothDenseLocalEvidence = lambda le, w, excludeMask=None, eps=1e-12: calls.append(int(w)) or "reference"
    cb.install(fake, background=False, munc=True)
    try:
        ev = np.ones((2, 64), np.float32)
        with pytest.warns(native.HostPathWarning, match="windowIntervals 9000"):
            assert fake.cMuncSmoothDenseLocalEvidence(ev, 9000) == "reference"
        assert calls == [9000]
    finally:
        cb.uninstall(fake)
    assert fake.cMuncSmoothDenseLocalEvidence(np.ones((2, 4), np.float32), 3) == "reference"
