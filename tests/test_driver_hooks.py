"""Driver-side reductions with a device version: ``core._relativeSignChangePerKB``.

CPU: the oracle restatement against golden values of the reference's own function and against that
function itself when oracle/_ref/driver is present.  GPU: the device half bit-exact against the oracle,
and the hook installed into the reference driver."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from golden.make_driver_golden import sign_change_inputs

DRV = os.path.join(ROOT, "oracle", "_ref", "driver")


def golden_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "driver_golden.npz"), allow_pickle=False)
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def ref_core():
    if not os.path.isdir(os.path.join(DRV, "consenrich")):
        pytest.skip("oracle/_ref/driver not built")
    if DRV not in sys.path:
        sys.path.insert(0, DRV)
    import consenrich.core as core
    return core


def test_oracle_matches_golden_values_of_the_reference_function(oracle):
    cases = golden_cases()
    assert len(cases) >= 4
    for name, c in cases.items():
        got = oracle.relativeSignChangePerKB(c["state"], c["data"], c["munc"], intervalSizeBP=25,
                                             background=c.get("background"), pad=1e-4)
        want = float(c["value"])
        assert (got is None and np.isnan(want)) or got == want, name


def test_oracle_matches_the_reference_function_on_fresh_seeds(oracle):
    core = ref_core()
    rng = np.random.default_rng(3)
    for m, n in ((1, 2), (2, 17), (5, 4000), (10, 30_001)):
        state, data, munc, bg = sign_change_inputs(rng, m, n)
        for b in (None, bg):
            want = core._relativeSignChangePerKB(state, data, munc, intervalSizeBP=50, background=b, pad=1e-4)
            assert oracle.relativeSignChangePerKB(state, data, munc, intervalSizeBP=50, background=b, pad=1e-4) == want


@pytest.mark.gpu
def test_device_residual_matches_oracle_bitwise(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from consenrich_b200 import driver
    rng = np.random.default_rng(8)
    for m, n in ((1, 1), (3, 33), (10, 70_001), (130, 3000), (10, 300_001)):
        state, data, munc, bg = sign_change_inputs(rng, m, n)
        for b in (None, bg):
            want = oracle.weighted_mean_residual(state, data, munc, b, 1e-4)
            got = driver.weighted_mean_residual(state, data, munc, b, 1e-4)
            np.testing.assert_array_equal(got, want)  # NaN positions included
    with pytest.raises(TypeError, match="float32 matrices"):
        driver.weighted_mean_residual(np.zeros(4), np.zeros((2, 4)), np.zeros((2, 4)))


@pytest.mark.gpu
def test_hook_installed_into_the_reference_driver():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    core = ref_core()
    from consenrich_b200 import driver
    rng = np.random.default_rng(9)
    state, data, munc, bg = sign_change_inputs(rng, 6, 50_000)
    want = [core._relativeSignChangePerKB(state, data, munc, intervalSizeBP=25, background=b, pad=1e-4) for b in (None, bg)]
    want64 = core._relativeSignChangePerKB(state, data.astype(np.float64), munc.astype(np.float64), intervalSizeBP=25)
    original = core._relativeSignChangePerKB
    driver.install_driver(core)
    try:
        assert core._relativeSignChangePerKB is not original
        got = [core._relativeSignChangePerKB(state, data, munc, intervalSizeBP=25, background=b, pad=1e-4) for b in (None, bg)]
        # float64 matrices are not covered on the device: they go to the function that was replaced, loudly
        with pytest.warns(driver.HostPathWarning):
            got64 = core._relativeSignChangePerKB(state, data.astype(np.float64), munc.astype(np.float64), intervalSizeBP=25)
        assert core._relativeSignChangePerKB(None, data, munc, intervalSizeBP=25) is None
    finally:
        driver.uninstall_driver(core)
    assert core._relativeSignChangePerKB is original
    assert got == want and got64 == want64


# ---- core._perIntervalOutputDiagnosticTracks (core.py:7734-7880) ----
def diag_inputs(rng, m, n, dim, variant):
    covar = np.zeros((n, 2, 2), np.float32)
    covar[:, 0, 0] = rng.uniform(0.01, 1.0, n)
    covar[:, 1, 1] = rng.uniform(0.001, 0.1, n)
    covar[:, 0, 1] = covar[:, 1, 0] = rng.uniform(-0.005, 0.005, n)
    if dim == 1:
        covar = np.ascontiguousarray(covar[:, :1, :1])
    munc = rng.uniform(0.05, 1.0, (m, n)).astype(np.float32)
    munc[rng.random((m, n)) < 0.02] = np.float32(1e30)
    kw = dict(stateCovarForward=covar, matrixMunc=munc, matrixQ0=np.diag([2e-3, 4e-4]).astype(np.float32),
              matrixF=np.array([[1, 1], [0, 1]], np.float32), stateCovarInit=1000.0,
              stateModel="level" if dim == 1 else "level_trend", lambdaExp=None, processPrecExp=None, processQScale=None,
              pNoiseForward=None, pad=1e-4, obsPrecisionMultiplierMin=0.1, obsPrecisionMultiplierMax=10.0,
              procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3)
    if variant in ("kappa", "all"):
        kw["processPrecExp"] = np.exp(rng.normal(0, 1.5, n)).astype(np.float32)
    if variant in ("lambda", "all"):
        kw["lambdaExp"] = rng.uniform(0.05, 20.0, n).astype(np.float32)
    if variant in ("qscale", "all"):
        qs = rng.uniform(0.5, 2.0, n).astype(np.float32)
        qs[0] = 1.0
        kw["processQScale"] = qs
    if variant == "pnoise":
        pn = (covar * 0.01).astype(np.float32)
        if n > 20:
            pn[7] = np.nan  # falls back to Q0 * qScale for interval 8
        kw["pNoiseForward"] = pn
    return kw


def diag_state_model(core, dim):
    return core.STATE_MODEL_LEVEL if dim == 1 else [getattr(core, k) for k in dir(core)
                                                    if k.startswith("STATE_MODEL_") and k != "STATE_MODEL_LEVEL"
                                                    and isinstance(getattr(core, k), str)][0]


def test_oracle_interval_diagnostics_match_the_reference_function(oracle):
    core = ref_core()
    rng = np.random.default_rng(5)
    for dim in (2, 1):
        for variant in ("plain", "kappa", "lambda", "qscale", "pnoise", "all"):
            kw = diag_inputs(rng, 3, 400, dim, variant)
            kw["stateModel"] = diag_state_model(core, dim)
            want = core._perIntervalOutputDiagnosticTracks(**kw)
            n = kw["stateCovarForward"].shape[0]
            obs = np.ones(n) if kw["lambdaExp"] is None else np.clip(np.asarray(kw["lambdaExp"], np.float64), 0.1, 10.0)
            pp = None if kw["processPrecExp"] is None else np.clip(np.asarray(kw["processPrecExp"], np.float64), 5e-3, 5e3)
            pn = None if (kw["pNoiseForward"] is None or pp is not None) else kw["pNoiseForward"]
            tr, _, g0, g1 = oracle.interval_diagnostics(kw["stateCovarForward"], kw["matrixMunc"], obs,
                                                        np.asarray(want["processQScale"], np.float64), pp, pn,
                                                        np.asarray(kw["matrixQ0"], np.float64), kw["matrixF"], dim, 1000.0, 1e-4)
            for name, got in (("muncTrace", tr), ("sumGain0", g0), ("sumGain1", g1)):
                np.testing.assert_allclose(got.astype(np.float32), want[name], rtol=3e-7, atol=0, err_msg=f"{dim} {variant} {name}")


@pytest.mark.gpu
def test_device_interval_diagnostics_match_oracle_bitwise(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from consenrich_b200 import driver
    rng = np.random.default_rng(6)
    for dim in (2, 1):
        for m, n in ((1, 1), (3, 33), (10, 70_001), (6, 600_001)):
            for variant in ("plain", "kappa", "pnoise", "all"):
                kw = diag_inputs(rng, m, n, dim, variant)
                obs = np.ones(n) if kw["lambdaExp"] is None else np.clip(np.asarray(kw["lambdaExp"], np.float64), 0.1, 10.0)
                qs = np.ones(n) if kw["processQScale"] is None else np.asarray(kw["processQScale"], np.float64)
                pp = None if kw["processPrecExp"] is None else np.clip(np.asarray(kw["processPrecExp"], np.float64), 5e-3, 5e3)
                pn = None if (kw["pNoiseForward"] is None or pp is not None) else kw["pNoiseForward"]
                args = (kw["stateCovarForward"], kw["matrixMunc"], obs, qs, pp, pn, np.asarray(kw["matrixQ0"], np.float64),
                        kw["matrixF"], dim, 1000.0, 1e-4)
                for got, want in zip(driver.interval_diagnostics(*args), oracle.interval_diagnostics(*args)):
                    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_interval_diagnostics_hook_in_the_reference_driver():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    core = ref_core()
    from consenrich_b200 import driver
    rng = np.random.default_rng(7)
    cases = []
    for dim in (2, 1):
        for variant in ("plain", "kappa", "lambda", "qscale", "pnoise", "all"):
            kw = diag_inputs(rng, 5, 20_000, dim, variant)
            kw["stateModel"] = diag_state_model(core, dim)
            cases.append((kw, core._perIntervalOutputDiagnosticTracks(**kw)))
    driver.install_driver(core)
    try:
        for kw, want in cases:
            got = core._perIntervalOutputDiagnosticTracks(**kw)
            assert list(got) == list(want)
            for name in want:
                assert got[name].dtype == np.float32 and got[name].shape == want[name].shape
                # float32 outputs of float64 arithmetic: equal up to the last float32 bit (numpy's 2x2 matrix
                # product may fuse multiply-adds; the kernel rounds every operation)
                np.testing.assert_allclose(got[name], want[name], rtol=3e-7, atol=0, err_msg=name)
        with pytest.raises(ValueError, match="matrixMunc must have shape"):
            core._perIntervalOutputDiagnosticTracks(**{**cases[0][0], "matrixMunc": cases[0][0]["matrixMunc"][:, :-1]})
    finally:
        driver.uninstall_driver(core)


def test_hooks_pass_uncovered_inputs_to_the_functions_they_replace():
    """No GPU needed: float64 matrices are not covered on the device, so the installed hooks hand them to the
    reference's own functions -- with a HostPathWarning, never silently; uninstall restores the originals."""
    core = ref_core()
    from consenrich_b200 import driver
    rng = np.random.default_rng(10)
    state, data, munc, bg = sign_change_inputs(rng, 3, 500)
    kw = diag_inputs(rng, 3, 300, 2, "all")
    kw["stateModel"] = diag_state_model(core, 2)
    kw64 = {**kw, "stateCovarForward": kw["stateCovarForward"].astype(np.float64)}
    originals = (core._relativeSignChangePerKB, core._perIntervalOutputDiagnosticTracks)
    want_sign = core._relativeSignChangePerKB(state, data.astype(np.float64), munc.astype(np.float64), intervalSizeBP=25,
                                              background=bg, pad=1e-4)
    want_diag = core._perIntervalOutputDiagnosticTracks(**kw64)
    driver.install_driver(core)
    try:
        assert core._relativeSignChangePerKB is not originals[0]
        assert core._perIntervalOutputDiagnosticTracks is not originals[1]
        with pytest.warns(driver.HostPathWarning, match="_relativeSignChangePerKB"):
            got_sign = core._relativeSignChangePerKB(state, data.astype(np.float64), munc.astype(np.float64),
                                                     intervalSizeBP=25, background=bg, pad=1e-4)
        with pytest.warns(driver.HostPathWarning, match="_perIntervalOutputDiagnosticTracks"):
            got_diag = core._perIntervalOutputDiagnosticTracks(**kw64)
        assert core._relativeSignChangePerKB(state, None, munc, intervalSizeBP=25) is None
    finally:
        driver.uninstall_driver(core)
    assert (core._relativeSignChangePerKB, core._perIntervalOutputDiagnosticTracks) == originals
    assert got_sign == want_sign
    for k in want_diag:
        np.testing.assert_array_equal(got_diag[k], want_diag[k])
