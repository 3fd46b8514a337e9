"""The reference's own test cases of the path, run with the ORACLE's restatements swapped into the reference's
extension module (CPU).  The oracle is pinned bit-exact to the reference build elsewhere (test_oracle*.py,
test_background_oracle.py, test_munc_oracle.py); this adds the reference's known-answer and contract cases on
top -- the same cases tests/test_reference_cases_gpu.py runs against the installed B200 kernels, so a case that
is green here and red there points at the device code or its Python boundary, not at the case."""
import sys

import pytest

from conftest import ROOT
from test_reference_cases_gpu import CASES, MUNC_CASES, MUNC_FUNCTIONS, _call, count_calls, ref_ns  # noqa: F401

HOT_PATH = ("cforwardPass", "cforwardPassLevel", "cbackwardPass", "cbackwardPassLevel", "cfixedBackgroundECM",
            "cfixedBackgroundECMLevel")
BACKGROUND = ("cbackgroundWeightedStats", "cbackgroundWeightedStatsWithSupport", "csolveZeroCenteredBackground")


@pytest.fixture(scope="module")
def oracle():
    sys.path.insert(0, ROOT)
    from oracle import oracle as module
    module.build()
    return module


def _swap(module, oracle, names):
    before = {k: getattr(module, k) for k in names}
    for k in names:
        setattr(module, k, getattr(oracle, k))

    def undo():
        for k, fn in before.items():
            setattr(module, k, fn)
    return undo


@pytest.mark.parametrize("name", list(CASES) + list(MUNC_CASES))
def test_reference_case_passes_on_the_oracle(ref_ns, oracle, name):  # noqa: F811
    how = CASES.get(name, MUNC_CASES.get(name))
    names = HOT_PATH + BACKGROUND + MUNC_FUNCTIONS
    if name not in ref_ns:
        pytest.skip(f"{name} is not defined in this reference checkout")
    try:
        _call(ref_ns, name, how)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"{name} fails on the reference build itself ({type(e).__name__}): drifted")
    mod = ref_ns["cconsenrich"]
    undo_swap = _swap(mod, oracle, names)
    counts, undo_count = count_calls(mod, names)
    try:
        _call(ref_ns, name, how)
    finally:
        undo_count()
        undo_swap()
    if name in MUNC_CASES or "RunConsenrich" in name or "CFixedBackground" in name or "ForwardBackward" in name:
        assert sum(counts.values()) > 0, "the case did not reach the swapped functions"
