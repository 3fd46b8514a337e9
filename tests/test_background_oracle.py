"""Pins the CPU restatement of the background-track functions (oracle/background_oracle.c):
(a) bit-exact against golden vectors produced by the reference build
    (tests/golden/make_background_golden.py), (b) bit-exact against the reference build itself when
    oracle/_ref is present, error cases included, (c) against a dense float64 solve of the same
    penalised system (an independent known answer)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR


def golden_cases(kind):
    z = np.load(os.path.join(GOLDEN_DIR, "background_golden.npz"), allow_pickle=False)
    cases = {}
    for key in z.files:
        k, name, field = key.split("/")
        if k == kind:
            cases.setdefault(name, {})[field] = z[key]
    return cases


def dense_system(w, lam, lam1):
    """diag(w) + lam1 D1'D1 + lam D2'D2 as a dense matrix (cconsenrich.pyx:905-943 says the same entries)."""
    n = w.shape[0]
    A = np.diag(w.astype(np.float64))
    if n >= 2 and lam1 > 0:
        D1 = np.diff(np.eye(n), axis=0)
        A += lam1 * D1.T @ D1
    if n >= 3 and lam > 0:
        D2 = np.diff(np.eye(n), n=2, axis=0)
        A += lam * D2.T @ D2
    return A


def test_oracle_matches_golden_background_vectors_bitwise(oracle):
    solve = golden_cases("solve")
    assert len(solve) >= 9
    for name, c in solve.items():
        got = oracle.csolveZeroCenteredBackground(c["weight"], c["rhs"], float(c["lam"]), bool(c["zeroCenter"]),
                                                  lamFirst=float(c["lamFirst"]))
        np.testing.assert_array_equal(got, c["out"], err_msg=name)
    for name, c in golden_cases("stats").items():
        w, r, sup = oracle.cbackgroundWeightedStatsWithSupport(c["resid"], c["inv"])
        np.testing.assert_array_equal(w, c["weight"], err_msg=name)
        np.testing.assert_array_equal(r, c["rhs"], err_msg=name)
        assert sup == int(c["support"])


def test_oracle_matches_reference_build_on_fresh_seeds(oracle):
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(77)
    for n in (1, 2, 3, 4, 6, 33, 500, 20001):
        for zc in (True, False):
            for lam, lam1 in ((128.0, 0.0), (2.5, 0.75), (0.0, 3.0), (0.0, 0.0)):
                w = rng.uniform(0.5, 80.0, n)
                w[rng.random(n) < 0.1] = 0.0
                if (lam, lam1) == (0.0, 0.0) and n % 2:
                    w = np.maximum(w, 0.5)
                rhs = rng.normal(size=n) * (w + 1.0)
                outs = []
                for mod in (ref, oracle):
                    try:
                        outs.append(("ok", mod.csolveZeroCenteredBackground(w, rhs, lam, zc, lamFirst=lam1)))
                    except (RuntimeError, ValueError) as e:
                        outs.append((type(e).__name__, str(e)))
                assert outs[0][0] == outs[1][0], (n, zc, lam, lam1, outs)
                if outs[0][0] == "ok":
                    np.testing.assert_array_equal(outs[0][1], outs[1][1])
                else:
                    assert outs[0][1] == outs[1][1]
    for m, n in ((1, 1), (4, 333), (12, 4097)):
        res = rng.normal(size=(m, n)).astype(np.float32)
        inv = rng.uniform(0, 5, (m, n)).astype(np.float32)
        inv[:, ::5] = 0
        a, b = ref.cbackgroundWeightedStatsWithSupport(res, inv), oracle.cbackgroundWeightedStatsWithSupport(res, inv)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        assert a[2] == b[2]
        a2, b2 = ref.cbackgroundWeightedStats(res, inv), oracle.cbackgroundWeightedStats(res, inv)
        np.testing.assert_array_equal(a2[0], b2[0])
        np.testing.assert_array_equal(a2[1], b2[1])


def test_oracle_solves_the_penalised_system(oracle):
    """Known answer: dense float64 solve; with zeroCenter the KKT system [A 1; 1' 0]."""
    rng = np.random.default_rng(3)
    for n, lam, lam1 in ((2, 4.0, 1.0), (3, 4.0, 0.0), (9, 50.0, 0.5), (200, 128.0, 0.0), (301, 0.0, 5.0)):
        w = rng.uniform(1.0, 30.0, n)
        rhs = rng.normal(size=n) * w
        A = dense_system(w, lam, lam1)
        x = oracle.csolveZeroCenteredBackground(w, rhs, lam, False, lamFirst=lam1)
        np.testing.assert_allclose(x, np.linalg.solve(A, rhs), rtol=0, atol=1e-10 * np.abs(x).max())
        xz = oracle.csolveZeroCenteredBackground(w, rhs, lam, True, lamFirst=lam1)
        K = np.block([[A, np.ones((n, 1))], [np.ones((1, n)), np.zeros((1, 1))]])
        want = np.linalg.solve(K, np.concatenate([rhs, [0.0]]))[:n]
        np.testing.assert_allclose(xz, want, rtol=0, atol=1e-10 * max(np.abs(want).max(), 1e-3))
        assert abs(xz.sum()) < 1e-9 * np.abs(xz).sum()


def test_oracle_argument_errors_carry_the_reference_texts(oracle):
    w = np.ones(4)
    with pytest.raises(ValueError, match="weightTrack and rhsTrack must have the same length"):
        oracle.csolveZeroCenteredBackground(w, np.ones(3), 1.0)
    with pytest.raises(ValueError, match="lam must be finite and nonnegative"):
        oracle.csolveZeroCenteredBackground(w, w, -1.0)
    with pytest.raises(ValueError, match="lamFirst must be finite and nonnegative"):
        oracle.csolveZeroCenteredBackground(w, w, 1.0, True, lamFirst=float("nan"))
    with pytest.raises(RuntimeError, match="pivot modification at index 0"):
        oracle.csolveZeroCenteredBackground(np.zeros(5), np.ones(5), 0.0, False)
    with pytest.raises(ValueError, match="identical 2D shapes"):
        oracle.cbackgroundWeightedStats(np.zeros((2, 3), np.float32), np.zeros((2, 4), np.float32))
    assert oracle.csolveZeroCenteredBackground(np.zeros(0), np.zeros(0), 1.0).shape == (0,)
