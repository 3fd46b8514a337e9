"""The reference's OWN hot-path test cases, run against the installed B200 kernels.

`/root/reference/tests/test_core.py` holds the known-answer and contract cases that pin this path
(SURVEY 8c): collected ones (`test_core_em_loop_contracts`, tests/test_core.py:7585) and dormant ones that
nothing calls but that pass on the reference build.  oracle/build_ref_driver.sh copies that file, unmodified,
next to the reference driver install (oracle/_ref/driver/ref_tests/, git-ignored, travels to the GPU box); it
imports modules that are not part of the driver install (peaks, detrorm, the htslib extension), so it is not
imported here: its top-level function definitions are exec'd into a namespace that binds the driver
install's `consenrich.core` / `consenrich.cconsenrich`, and the named `_case...` functions are called --
first on the reference's own Cython kernels (a case that fails THERE has drifted and is skipped), then with
`consenrich_b200.install()` active.  Adaptive process noise included (`_caseRunConsenrichAPNSmoke`)."""
import ast
import logging
import math
import os
import sys
import tempfile
import types

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver")
SOURCE = os.path.join(DRIVER, "ref_tests", "test_core_source.py")

# name -> False: called without arguments; True: called with a MonkeyPatch; a list of tuples: called once per tuple
CASES = {
    # collected by test_core_em_loop_contracts (tests/test_core.py:7585-7607)
    "_caseRunConsenrichOuterPassSmoke": False,
    "_caseRunConsenrichInitialProcessQSkipsWarmup": True,
    "_caseRunConsenrichFixedDiagonalUsesDataQ": True,
    "_caseRunConsenrichOuterPassRequiresThreeIterationsDespiteTolerance": True,
    "_caseRunConsenrichAlwaysRunsECMWithAPN": True,
    "_caseRunConsenrichAPNSmoke": False,
    "_caseRunConsenrichLevelStateModelSmoke": False,
    # dormant known-answer cases of the kernels themselves
    "_caseLevelForwardBackwardMatchesPythonReference": False,
    "_caseLevelEmbeddedForwardBackwardAgreementWithPrecisionMultipliers": False,
    "_caseCFixedBackgroundPrecisionUpdatesMatchStudentTEquations": False,
    "_caseCFixedBackgroundECMTinyTrackUsesFiniteFallback": False,
    "_caseCFixedBackgroundECMLevelTinyTrackUsesFiniteFallback": False,
    "_caseObservationPrecisionIsIntervalLevelOnly": False,
    "_caseExpectedTransitionResidualSumsUsesLagOrientationAndDeltaF": False,
    "_caseExpectedTransitionResidualSumsMatchesPythonReference": False,
    "_caseFinalForwardNISUsesMeanFinalForwardDiagnostic": False,
    "_caseFinalForwardGainSummaryUsesReplicateContigRows": False,
    "_casePerIntervalOutputDiagnosticsUseEffectiveNoiseAndGainComponents": False,
}

# the observation-noise stage: the reference's cases of its four compiled functions, alone and under its own
# `core.getMuncTrack` (core.py:8390), which reaches them by attribute (core.py:8526-8830)
MUNC_FUNCTIONS = ("cMuncSmoothDenseLocalEvidence", "cFinalizeMuncEBTrack", "cMuncObservationMomentSeedPass", "cEMA")
MUNC_CASES = {
    "_caseCEMAUsesSameBidirectionalKernelForFloat32AndFloat64": [(np.float32,), (np.float64,)],
    "_caseMuncObservationMomentSeedPassUsesOmegaMomentsAndFloors": False,
    "_caseFinalizeMuncEBTrackPreservesCountFloorSentinel": False,
    "_caseMuncSmoothDenseLocalEvidenceUsesCenteredWindows": False,
    "_caseGetMuncTrackAppliesAdditiveCovariatesBeforeEBShrinkage": False,
    "_caseGetMuncTrackRejectsSparseLocalVariancePaths": False,
    "_caseGetMuncTrackCapsPriorStrengthAtFiftyTimesLocalDf": True,
    "_caseGetMuncTrackSmoothsPriorMeanWithEMA": True,
    "_caseGetMuncTrackAppliesReplicateVarianceFactor": False,
}


class _Missing(types.ModuleType):
    """Stand-in for a reference module outside the driver install: any use fails with its name."""

    def __getattr__(self, item):
        raise RuntimeError(f"{self.__name__}.{item}: module not part of oracle/_ref/driver")


@pytest.fixture(scope="module")
def ref_ns():
    if not os.path.isfile(SOURCE):
        pytest.skip("oracle/_ref/driver/ref_tests not built (oracle/build_ref_driver.sh needs /root/reference)")
    sys.path.insert(0, DRIVER)
    import consenrich.cconsenrich as cconsenrich
    import consenrich.constants as constants
    import consenrich.core as core
    import consenrich.diagnostics as diagnostics
    import consenrich.misc_util as misc_util
    import pandas as pd
    import scipy.ndimage as ndi
    import scipy.signal as spySig
    import scipy.stats as stats
    from pathlib import Path
    from types import SimpleNamespace
    from typing import List, Optional, Tuple
    logging.getLogger("consenrich").setLevel(logging.ERROR)
    logging.disable(logging.WARNING)
    ns = dict(logging=logging, math=math, os=os, tempfile=tempfile, SimpleNamespace=SimpleNamespace, Tuple=Tuple, List=List,
              Optional=Optional, Path=Path, pd=pd, pytest=pytest, np=np, ndi=ndi, stats=stats, spySig=spySig, core=core,
              constants=constants, cconsenrich=cconsenrich, diagnostics=diagnostics, misc_util=misc_util,
              consenrichRuntime=_Missing("consenrich.consenrich"), ccounts=_Missing("consenrich.ccounts"),
              detrorm=_Missing("consenrich.detrorm"), peaks=_Missing("consenrich.peaks"), __name__="ref_test_core")
    tree = ast.parse(open(SOURCE).read(), SOURCE)
    # imports dropped (bound above); everything else at module level -- helper functions, constants,
    # the cases -- kept as written
    tree.body = [node for node in tree.body if not isinstance(node, (ast.Import, ast.ImportFrom))]
    kept = []
    for node in tree.body:
        mod = ast.Module(body=[node], type_ignores=[])
        try:
            exec(compile(mod, SOURCE, "exec"), ns)
            kept.append(node)
        except Exception:  # a module-level statement that needs one of the modules left out
            continue
    yield ns
    logging.disable(logging.NOTSET)
    sys.path.remove(DRIVER)


def _call(ns, name, how):
    fn = ns[name]
    if how is True:
        mp = pytest.MonkeyPatch()
        try:
            fn(mp)
        finally:
            mp.undo()
    elif how is False:
        fn()
    else:
        for args in how:
            fn(*args)


def count_calls(module, names):
    """Wrap ``module``'s attributes ``names`` with counters; returns (counts, undo)."""
    counts = {k: 0 for k in names}
    before = {k: getattr(module, k) for k in names}

    def wrap(k, fn):
        def counted(*args, **kwargs):
            counts[k] += 1
            return fn(*args, **kwargs)
        return counted
    for k in names:
        setattr(module, k, wrap(k, before[k]))

    def undo():
        for k in names:
            setattr(module, k, before[k])
    return counts, undo


@pytest.mark.parametrize("name", list(CASES))
def test_reference_case_passes_on_the_installed_kernels(ref_ns, name):
    import consenrich_b200 as cb
    if name not in ref_ns:
        pytest.skip(f"{name} is not defined in this reference checkout")
    try:
        _call(ref_ns, name, CASES[name])  # on the reference's own kernels first
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"{name} fails on the reference build itself ({type(e).__name__}): drifted, not a parity statement")
    mod = cb.install(ref_ns["cconsenrich"])
    cb.install_driver(ref_ns["core"])
    try:
        assert mod.cforwardPass is cb.cforwardPass and mod.cfixedBackgroundECM is cb.cfixedBackgroundECM
        launches0 = cb._lib.default_context().launch_count
        _call(ref_ns, name, CASES[name])
        if "RunConsenrich" in name or "CFixedBackground" in name or "ForwardBackward" in name:
            assert cb._lib.default_context().launch_count > launches0, "the case did not reach the device"
    finally:
        cb.uninstall_driver(ref_ns["core"])
        cb.uninstall(ref_ns["cconsenrich"])


@pytest.mark.parametrize("name", list(MUNC_CASES))
def test_reference_munc_case_passes_on_the_installed_kernels(ref_ns, name):
    import consenrich_b200 as cb
    if name not in ref_ns:
        pytest.skip(f"{name} is not defined in this reference checkout")
    try:
        _call(ref_ns, name, MUNC_CASES[name])
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"{name} fails on the reference build itself ({type(e).__name__}): drifted, not a parity statement")
    mod = cb.install(ref_ns["cconsenrich"], munc=True)
    try:
        assert mod.cEMA is cb.cEMA and mod.cFinalizeMuncEBTrack is cb.cFinalizeMuncEBTrack
        counts, undo = count_calls(mod, MUNC_FUNCTIONS)
        try:
            _call(ref_ns, name, MUNC_CASES[name])
        finally:
            undo()
        assert sum(counts.values()) > 0, "the case did not reach the installed functions"
    finally:
        cb.uninstall(ref_ns["cconsenrich"])
