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
        # float64 matrices are not covered on the device: they go to the function that was replaced
        got64 = core._relativeSignChangePerKB(state, data.astype(np.float64), munc.astype(np.float64), intervalSizeBP=25)
        assert core._relativeSignChangePerKB(None, data, munc, intervalSizeBP=25) is None
    finally:
        driver.uninstall_driver(core)
    assert core._relativeSignChangePerKB is original
    assert got == want and got64 == want64
