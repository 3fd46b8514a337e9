"""Pins the CPU restatement of the observation-noise stage's dense kernels (oracle/munc_oracle.c):
bit-exact against golden vectors of the reference build (tests/golden/make_munc_golden.py, the
reference's own known-answer case included) and against the reference build itself when present."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR


def golden_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "munc_golden.npz"), allow_pickle=False)
    cases = {}
    for key in z.files:
        _, name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


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
