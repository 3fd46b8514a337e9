"""Pins the CPU oracle (oracle/ssm_oracle.c + oracle/oracle.py).

(a) golden vectors produced by the reference build (tests/golden/make_golden.py) -- bit-exact;
(b) the reference build itself (oracle/_ref) when present -- bit-exact on fresh seeds;
(c) an independent float64 recursion of the kind the reference's own tests use as a
    known answer (reference tests/test_core.py:522-592, 3276-3351; tolerance 2e-6).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, synth_tracks

F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)


def _golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
    cases = {}
    for key in z.files:
        case, rest = key.split("/", 1)
        cases.setdefault(case, {})[rest] = z[key]
    return cases


def _run_sweep(mod, case, dim):
    data, munc, Q0, bm = case["data"], case["munc"], case["Q0"], case["blockMap"]
    n = data.shape[1]
    extra = {}
    for k, v in case.items():
        if k.startswith("extra/"):
            v = v if v.ndim else v.item()
            extra[k[6:]] = v
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=bm,
              blockCount=int(bm.max()) + 1, stateInit=float(case["stateInit"]), stateCovarInit=1000.0,
              pad=1.0e-4, returnNLL=True, **extra)
    st = dict(stateForward=np.empty((n, dim), np.float32), stateCovarForward=np.empty((n, dim, dim), np.float32),
              pNoiseForward=np.zeros((n, dim, dim), np.float32), vectorD=np.empty(n, np.float32))
    if dim == 2:
        r = mod.cforwardPass(matrixF=F, **kw, **st)
        b = mod.cbackwardPass(matrixData=data, matrixF=F, stateForward=st["stateForward"],
                              stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"])
    else:
        r = mod.cforwardPassLevel(**kw, **st)
        b = mod.cbackwardPassLevel(matrixData=data, stateForward=st["stateForward"],
                                   stateCovarForward=st["stateCovarForward"], pNoiseForward=st["pNoiseForward"])
    out = dict(phiHat=np.float32(r[0]), sumNLL=np.float64(r[3]), **st)
    out.update(zip(("stateSmoothed", "stateCovarSmoothed", "lagCovSmoothed", "postFitResiduals"), b))
    return out


@pytest.mark.parametrize("dim", [2, 1])
def test_oracle_matches_golden_sweeps_bitwise(oracle, dim):
    cases = _golden("sweep_golden.npz")
    assert len(cases) >= 6
    for name, case in cases.items():
        got = _run_sweep(oracle, case, dim)
        n = case["data"].shape[1]
        for key, val in got.items():
            want = case[f"d{dim}/{key}"]
            if key == "lagCovSmoothed" and n == 1:
                continue  # single uninitialised row in the reference (np.empty)
            np.testing.assert_array_equal(val, want, err_msg=f"{name} d{dim} {key}")


def _run_ecm(mod, case, dim):
    opts = {k[5:]: (v.item()) for k, v in case.items() if k.startswith("opts/")}
    n = case["data"].shape[1]
    kw = dict(matrixData=case["data"], matrixPluginMuncInit=case["munc"], matrixQ0=case["Q0"],
              intervalToBlockMap=np.zeros(n, np.int32), blockCount=1, stateInit=0.0, stateCovarInit=1000.0,
              returnIntermediates=True, returnDiagnostics=True, logIterations=False, **opts)
    return mod.cfixedBackgroundECM(matrixF=F, **kw) if dim == 2 else mod.cfixedBackgroundECMLevel(**kw)


@pytest.mark.parametrize("dim", [2, 1])
def test_oracle_matches_golden_ecm_bitwise(oracle, dim):
    cases = _golden("ecm_golden.npz")
    assert len(cases) >= 4
    for name, case in cases.items():
        out = _run_ecm(oracle, case, dim)
        pre = f"d{dim}/"
        assert out[0] == int(case[pre + "itersDone"]), name
        assert out[1] == float(case[pre + "nll"]), name
        assert out[8]["converged"] == bool(case[pre + "converged"])
        for nm, arr in zip(("stateSmoothed", "stateCovarSmoothed", "lagCovSmoothed", "postFitResiduals",
                            "lambdaExp", "processPrecExp"), out[2:8]):
            if arr is None:
                assert (pre + nm) not in case
            else:
                np.testing.assert_array_equal(arr, case[pre + nm], err_msg=f"{name} d{dim} {nm}")


def test_oracle_matches_reference_build_bitwise(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here (golden vectors still pin the oracle)")
    rng = np.random.default_rng(7)
    for seed, (m, n) in enumerate([(2, 33), (8, 4097), (25, 700)]):
        data, munc = synth_tracks(100 + seed, m, n, masked_frac=0.03)
        case = dict(data=data, munc=munc, Q0=np.array([[1e-3, 0], [0, 2e-4]], np.float32),
                    blockMap=np.zeros(n, np.int32), stateInit=np.float32(0.1))
        qs = (0.5 + rng.random(n)).astype(np.float32)
        qs[0] = 1.0
        case["extra/lambdaExp"] = (0.1 + 5 * rng.random(n)).astype(np.float32)
        case["extra/processPrecExp"] = np.exp(rng.normal(0, 2, n)).astype(np.float32)
        case["extra/processQScale"] = qs
        for dim in (2, 1):
            a, b = _run_sweep(oracle, case, dim), _run_sweep(ref, case, dim)
            for key in a:
                np.testing.assert_array_equal(a[key], b[key], err_msg=f"{m}x{n} d{dim} {key}")
        ecase = dict(data=data, munc=munc, Q0=case["Q0"])
        ecase["opts/ECM_fixedBackgroundIters"] = np.asarray(3)
        ecase["opts/ECM_fixedBackgroundRtol"] = np.asarray(1e-6)
        for dim in (2, 1):
            a, b = _run_ecm(oracle, ecase, dim), _run_ecm(ref, ecase, dim)
            assert a[0] == b[0] and a[1] == b[1]
            for x, y in zip(a[2:8], b[2:8]):
                np.testing.assert_array_equal(x, y)
            assert a[8] == b[8]


def _float64_level_reference(data, munc, q, x0, p0, pad):
    """Independent float64 scalar Kalman filter + RTS smoother in textbook (gain) form."""
    z, v = data.astype(np.float64), np.maximum(munc.astype(np.float64) + pad, 1e-12)
    m, n = z.shape
    xf, pf = np.empty(n), np.empty(n)
    x, p = float(x0), float(p0)
    for k in range(n):
        p = p + q
        prec = (1.0 / v[:, k]).sum()
        ybar = (z[:, k] / v[:, k]).sum() / prec
        gain = p * prec / (1.0 + p * prec)
        x = x + gain * (ybar - x)
        p = (1.0 - gain) * p
        xf[k], pf[k] = x, p
    xs, ps, lag = xf.copy(), pf.copy(), np.empty(max(n - 1, 1))
    for k in range(n - 2, -1, -1):
        pp = pf[k] + q
        g = pf[k] / pp
        xs[k] = xf[k] + g * (xs[k + 1] - xf[k])
        ps[k] = pf[k] + g * g * (ps[k + 1] - pp)
        lag[k] = g * ps[k + 1]
    return xf, pf, xs, ps, lag, z.T - xs[:, None]


def test_oracle_level_model_known_answer(oracle):
    data, munc = synth_tracks(3, 2, 64)
    n, q = data.shape[1], 0.06
    st = dict(stateForward=np.empty((n, 1), np.float32), stateCovarForward=np.empty((n, 1, 1), np.float32),
              pNoiseForward=np.empty((n, 1, 1), np.float32))
    oracle.cforwardPassLevel(matrixData=data, matrixPluginMuncInit=munc,
                             matrixQ0=np.array([[q, 0], [0, 0.5]], np.float32),
                             intervalToBlockMap=np.zeros(n, np.int32), blockCount=1, stateInit=-0.1,
                             stateCovarInit=0.8, pad=0.02, ECM_useObsPrecisionReweighting=False,
                             ECM_useProcessPrecisionReweighting=False, **st)
    xs, ps, lag, res = oracle.cbackwardPassLevel(matrixData=data, **st)
    q32, pad32 = float(np.float32(q)), float(np.float32(0.02))
    rxf, rpf, rxs, rps, rlag, rres = _float64_level_reference(
        data, munc, q32, float(np.float32(-0.1)), float(np.float32(0.8)), pad32)
    tol = dict(rtol=2e-6, atol=2e-6)  # the reference's own tolerance, tests/test_core.py:3344-3351
    np.testing.assert_allclose(st["stateForward"][:, 0], rxf, **tol)
    np.testing.assert_allclose(st["stateCovarForward"][:, 0, 0], rpf, **tol)
    np.testing.assert_allclose(xs[:, 0], rxs, **tol)
    np.testing.assert_allclose(ps[:, 0, 0], rps, **tol)
    np.testing.assert_allclose(lag[: n - 1, 0, 0], rlag[: n - 1], **tol)
    np.testing.assert_allclose(res, rres, **tol)


def test_oracle_two_state_identity_transition_embeds_level_model(oracle):
    """reference tests/test_core.py:3355-3470: with F = I the 2-state level equals the 1-state model."""
    data, munc = synth_tracks(5, 3, 90)
    n = data.shape[1]
    rng = np.random.default_rng(11)
    qs = (0.5 + rng.random(n)).astype(np.float32)
    qs[0] = 1.0
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, intervalToBlockMap=np.zeros(n, np.int32), blockCount=1,
              stateInit=-0.15, stateCovarInit=0.7, pad=0.015, returnNLL=True, storeNLLInD=True,
              lambdaExp=(0.1 + 7 * rng.random(n)).astype(np.float32),
              processPrecExp=(0.2 + 6 * rng.random(n)).astype(np.float32), processQScale=qs)
    Q0 = np.array([[0.045, 0.0], [0.0, 0.125]], np.float32)
    s1 = dict(stateForward=np.empty((n, 1), np.float32), stateCovarForward=np.empty((n, 1, 1), np.float32),
              pNoiseForward=np.empty((n, 1, 1), np.float32), vectorD=np.empty(n, np.float32))
    s2 = dict(stateForward=np.empty((n, 2), np.float32), stateCovarForward=np.empty((n, 2, 2), np.float32),
              pNoiseForward=np.empty((n, 2, 2), np.float32), vectorD=np.empty(n, np.float32))
    r1 = oracle.cforwardPassLevel(matrixQ0=Q0, **kw, **s1)
    r2 = oracle.cforwardPass(matrixF=np.eye(2, dtype=np.float32), matrixQ0=Q0, **kw, **s2)
    assert r1[2] is s1["vectorD"] and r2[2] is s2["vectorD"]
    assert r2[3] == pytest.approx(r1[3], rel=2e-6, abs=2e-6)
    tol = dict(rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(s2["vectorD"], s1["vectorD"], **tol)
    np.testing.assert_allclose(s2["stateForward"][:, :1], s1["stateForward"], **tol)
    np.testing.assert_allclose(s2["stateCovarForward"][:, :1, :1], s1["stateCovarForward"], **tol)


def test_oracle_error_conventions(oracle):
    data, munc = synth_tracks(1, 2, 10)
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixF=F, matrixQ0=np.eye(2, dtype=np.float32) * 1e-3,
              intervalToBlockMap=np.zeros(10, np.int32), blockCount=1, stateInit=0.0, stateCovarInit=1.0)
    with pytest.raises(ValueError, match="blockCount must be positive"):
        oracle.cforwardPass(**{**kw, "blockCount": 0})
    with pytest.raises(ValueError, match="out-of-range block id"):
        oracle.cforwardPass(**{**kw, "intervalToBlockMap": np.full(10, 3, np.int32)})
    with pytest.raises(ValueError, match=r"processQScale\[0\] must be 1.0"):
        oracle.cforwardPass(**kw, processQScale=np.full(10, 2.0, np.float32))
    with pytest.raises(ValueError, match="singular"):
        oracle.cfixedBackgroundECM(**{**kw, "matrixQ0": np.zeros((2, 2), np.float32)}, logIterations=False)
    empty = np.empty((2, 0), np.float32)
    r = oracle.cforwardPass(**{**kw, "matrixData": empty, "matrixPluginMuncInit": empty}, returnNLL=True)
    assert r[0] == 0.0 and r[3] == 0.0 and r[2].shape == (0,)
