"""GPU parity of the ECM's lean inner sweeps (csrc/lean_kernels.cu: run-major private tracks, three
launches per forward pass, two per backward pass) against the CPU oracle and against the look-back
scan kernels on the same inputs, through the C ABI.

Covered: both run lengths (32 / 64 intervals), tracks that end inside a run / inside a segment / on
a segment boundary, masked cells, processQScale, a warm-started kappa, fixed budgets and the
free-running stopping rule, and the chr19-sized call the benchmark times.

Tolerance: tests/parity_util.py for lean-vs-oracle (the same stated tolerance as the sweeps); the two
device paths replay the same float32-rounded recursion from start states that differ only by float64
re-association, so they are compared 10x tighter.
"""
import numpy as np
import pytest

from conftest import synth_tracks
from parity_util import assert_sweep_tracks_close, assert_tracks_close

pytestmark = pytest.mark.gpu

F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
Q0 = np.array([[2e-3, 0.0], [0.0, 1e-4]], np.float32)
CLI = dict(ECM_useObsPrecisionReweighting=False, ECM_useProcessPrecisionReweighting=True,
           procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3)


@pytest.fixture(scope="module")
def cb():
    import consenrich_b200 as cb
    cb._lib.default_context(0)
    return cb


@pytest.fixture()
def lean(cb):
    """lean(mode, log2_run) switches the path; always restored to the default afterwards."""
    L = cb._lib.load()

    def set_(mode, log2_run=0):
        cb._lib.check(L.cb200_set_lean_sweeps(mode, log2_run))
    yield set_
    cb._lib.check(L.cb200_set_lean_sweeps(1, 0))


def _ecm(mod, data, munc, **opts):
    n = data.shape[1]
    return mod.cfixedBackgroundECM(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=Q0,
                                   intervalToBlockMap=np.zeros(n, np.int32), blockCount=1, stateInit=0.0,
                                   stateCovarInit=1000.0, returnIntermediates=True, returnDiagnostics=True,
                                   logIterations=False, **{**CLI, **opts})


def _close(a, b, label, tight=1.0):
    """a, b: 9-tuples of cfixedBackgroundECM.  tight < 1 shrinks the stated tolerance."""
    assert a[0] == b[0], label
    assert abs(a[1] - b[1]) <= 1e-7 * max(abs(b[1]), 1.0), f"{label}: nll {a[1]} vs {b[1]}"
    kw = dict(rtol=1e-4 * tight, atol_rel=1e-5 * tight)
    # the rows under the diffuse prior get the sweeps' transient allowance (tests/parity_util.py)
    assert_sweep_tracks_close(a[2], b[2], f"{label} stateSmoothed", **kw)
    assert_sweep_tracks_close(a[3], b[3], f"{label} stateCovarSmoothed", scale="component", **kw)
    assert_sweep_tracks_close(a[4], b[4], f"{label} lagCovSmoothed", scale="component", **kw)
    assert_sweep_tracks_close(a[5], b[5], f"{label} residuals", **kw)
    assert a[6] is None and b[6] is None
    assert_tracks_close(a[7], b[7], f"{label} kappa", rtol=2e-4 * tight, atol_rel=1e-5 * tight)


@pytest.mark.parametrize("log2_run", [5, 6])
@pytest.mark.parametrize("m,n", [(3, 4096), (5, 4097), (8, 6000), (2, 32 * 64 * 3), (7, 40_003), (4, 65_537),
                                 (3, 131_071)])
def test_lean_ecm_matches_oracle_and_lookback(cb, oracle, lean, log2_run, m, n):
    data, munc = synth_tracks(7000 + n, m, n, masked_frac=0.02)
    opts = dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, t_innerIters=2)
    want = _ecm(oracle, data, munc, **opts)
    lean(1, log2_run)
    got = _ecm(cb, data, munc, **opts)
    lean(0)
    ref = _ecm(cb, data, munc, **opts)
    _close(got, want, f"lean L=2^{log2_run} {m}x{n} vs oracle")
    _close(got, ref, f"lean L=2^{log2_run} {m}x{n} vs look-back", tight=0.1)
    # the level residual identity holds exactly on the published tracks
    lvl = got[2][:, 0].astype(np.float64)
    np.testing.assert_array_equal(got[5], (data.T.astype(np.float64) - lvl[:, None]).astype(np.float32))
    np.testing.assert_array_equal(got[3][:, 0, 1], got[3][:, 1, 0])
    assert got[7][0] == 1.0


def test_lean_ecm_qscale_warm_start_and_stopping_rule(cb, oracle, lean):
    m, n = 6, 50_001
    data, munc = synth_tracks(99, m, n, masked_frac=0.05)
    rng = np.random.default_rng(8)
    qs = (0.5 + rng.random(n)).astype(np.float32)
    qs[0] = 1.0
    warm = np.exp(rng.normal(0, 1, n)).astype(np.float32)
    opts = dict(ECM_fixedBackgroundIters=20, ECM_fixedBackgroundRtol=1e-5, t_innerIters=3, processQScale=qs,
                processPrecExpInit=warm, ECM_robustTNu=5.0)
    want = _ecm(oracle, data, munc, **opts)
    lean(1)
    got = _ecm(cb, data, munc, **opts)
    _close(got, want, "lean free-running")
    assert got[8]["converged"] == want[8]["converged"]
    assert 2 <= got[0] < 20  # stopped by the rule, not by the budget


def test_lean_ecm_is_deterministic(cb, lean):
    data, munc = synth_tracks(5, 4, 200_003)
    opts = dict(ECM_fixedBackgroundIters=2, ECM_fixedBackgroundRtol=0.0, t_innerIters=2)
    lean(1)
    a, b = _ecm(cb, data, munc, **opts), _ecm(cb, data, munc, **opts)
    assert a[1] == b[1]
    for i in (2, 3, 4, 5, 7):
        np.testing.assert_array_equal(a[i], b[i])  # no inter-CTA ordering anywhere on this path


def test_lean_ecm_benchmark_call(cb, oracle, lean):
    """The call bench.py times at cfg2 (10 x 2 344 705, K = 3, t = 5), on a slice the oracle finishes in
    seconds and at the full length against the look-back path."""
    m, n = 10, 2_344_705
    rng = np.random.default_rng(1729)
    k = np.arange(n, dtype=np.float32)
    x = (0.5 * np.sin(k / 8000.0) + 2.0 * (np.sin(k / 733.0) > 0.98)).astype(np.float32)
    munc = (0.15 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n)).astype(np.float32))
    data = (x[None, :] + rng.standard_normal((m, n), dtype=np.float32) * np.sqrt(munc)).astype(np.float32)
    opts = dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=5, ECM_robustTNu=8.0)
    ns = 300_000
    ds, vs = np.ascontiguousarray(data[:, :ns]), np.ascontiguousarray(munc[:, :ns])
    lean(1)
    _close(_ecm(cb, ds, vs, **opts), _ecm(oracle, ds, vs, **opts), "bench call, 300k slice")
    got = _ecm(cb, data, munc, **opts)
    lean(0)
    ref = _ecm(cb, data, munc, **opts)
    _close(got, ref, "bench call, full chr19, lean vs look-back", tight=0.1)
