"""GPU parity of cMuncSmoothDenseLocalEvidence through the C ABI.

The reference slides a float64 running sum along each row; the kernel reads each window as a
difference of tile-local float64 prefix sums.  The two sums differ by ~1e-13 relative, so the float32
outputs are identical except where that lands on a rounding boundary: stated tolerance 1 float32 ulp
(1.2e-7 relative), and all but a vanishing fraction of cells must be bit-equal."""
import numpy as np
import pytest

from golden.make_munc_golden import exclude_mask, local_evidence
from test_munc_oracle import golden_cases, run_case

pytestmark = pytest.mark.gpu
ULP = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def cb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import consenrich_b200 as cb
    return cb


def assert_one_ulp(got, want, what):
    assert got.shape == want.shape and got.dtype == np.float32, what
    g, w = got.astype(np.float64), want.astype(np.float64)
    assert np.all(np.abs(g - w) <= ULP * np.abs(w)), f"{what}: beyond one float32 ulp"
    differing = np.count_nonzero(got != want)
    assert differing <= max(2, 1e-5 * got.size), f"{what}: {differing} of {got.size} cells differ"


def test_matches_reference_golden_vectors(cb):
    for name, c in golden_cases().items():
        assert_one_ulp(run_case(cb, c), c["out"], name)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_matches_oracle_on_fresh_seeds(cb, oracle, mode):
    rng = np.random.default_rng(100 + mode)
    shapes = [(1, 1, 1), (2, 5, 3), (3, 2047, 4), (3, 2048, 5), (2, 2049, 64), (10, 70001, 41), (4, 9000, 2048),
              (2, 30000, 8192), (1, 100, 1000), (130, 3000, 16)]
    for m, n, w in shapes:
        le, mk = local_evidence(rng, m, n), exclude_mask(rng, m, n, mode)
        want = oracle.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=1e-6)
        got = cb.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=1e-6)
        assert_one_ulp(got, want, f"{m}x{n} window {w} mask mode {mode}")


def test_full_size_properties(cb):
    """hg38 chr19 @ 25 bp x 10 tracks: constant rows come back unchanged; a fully masked stretch returns
    the cells themselves; outputs are bounded by the window's extremes."""
    m, n, w = 10, 2_344_705, 41
    rng = np.random.default_rng(1)
    const = np.full((m, n), np.float32(0.375))
    np.testing.assert_array_equal(cb.cMuncSmoothDenseLocalEvidence(const, w), const)
    le = local_evidence(rng, m, n)
    mk = np.zeros(n, np.uint8)
    mk[1000:1200] = 1
    out = cb.cMuncSmoothDenseLocalEvidence(le, w, excludeMask=mk, eps=1e-12)
    np.testing.assert_array_equal(out[:, 1030:1170], le[:, 1030:1170])  # windows with no unmasked cell
    assert out.min() >= le.min() * (1 - 1e-6) and out.max() <= le.max() * (1 + 1e-6)
    j, i = 3, 1_500_000
    want = np.float32(le[j, i - w // 2: i - w // 2 + w].astype(np.float64).mean())
    assert abs(float(out[j, i]) - float(want)) <= ULP * float(want)


def test_errors_follow_the_reference(cb):
    le = np.ones((2, 40), np.float32)
    with pytest.raises(ValueError, match="windowIntervals must be positive"):
        cb.cMuncSmoothDenseLocalEvidence(le, 0)
    with pytest.raises(ValueError, match="eps must be positive and finite"):
        cb.cMuncSmoothDenseLocalEvidence(le, 3, eps=-1.0)
    with pytest.raises(ValueError, match="excludeMask length must match interval count"):
        cb.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=np.zeros(41, np.uint8))
    with pytest.raises(ValueError, match="excludeMask shape must match localEvidence shape"):
        cb.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=np.zeros((2, 41), np.uint8))
    with pytest.raises(ValueError, match="excludeMask must be one- or two-dimensional"):
        cb.cMuncSmoothDenseLocalEvidence(le, 3, excludeMask=np.zeros((1, 2, 40), np.uint8))
    bad = le.copy()
    bad[1, 7] = np.nan
    with pytest.raises(ValueError, match="active local evidence cells must be positive and finite"):
        cb.cMuncSmoothDenseLocalEvidence(bad, 5)
    mk = np.zeros(40, np.uint8)
    mk[7] = 1
    assert cb.cMuncSmoothDenseLocalEvidence(bad, 5, excludeMask=mk).shape == (2, 40)
    with pytest.raises(NotImplementedError, match="not supported on the device"):
        cb.cMuncSmoothDenseLocalEvidence(le, 8193)
    assert cb.cMuncSmoothDenseLocalEvidence(np.zeros((0, 5), np.float32), 3).shape == (0, 5)


# ---- cFinalizeMuncEBTrack: elementwise, float64 arithmetic with the reference's rounding sequence ----
def test_finalize_matches_golden_vectors_bitwise(cb):
    from test_munc_oracle import check_finalize, run_finalize
    for name, c in golden_cases("finalize").items():
        check_finalize(run_finalize(cb, c), c)


def test_finalize_matches_oracle_bitwise_on_fresh_seeds(cb, oracle):
    rng = np.random.default_rng(31)
    for n in (1, 31, 257, 70001, 2_344_705):
        loc = np.exp(rng.normal(-2.0, 2.0, n)).astype(np.float32)
        pri = np.exp(rng.normal(-2.0, 1.0, n)).astype(np.float32)
        cf = rng.uniform(0.0, 0.3, n).astype(np.float32)
        cf[rng.random(n) < 0.3] = np.nan
        cf[rng.random(n) < 0.2] = 0.0
        for kw in (dict(nuLocal=37.0, nuPrior=12.25, varianceFloor=1e-3, varianceCap=3.0),
                   dict(nuLocal=5.0, nuPrior=1e-3), dict(useEB=False, varianceFloor=0.02, varianceCap=0.5)):
            for use_cf in (True, False):
                args = (loc, None if kw.get("useEB") is False else pri, cf if use_cf else None)
                want = oracle.cFinalizeMuncEBTrack(*args, **kw)
                got = cb.cFinalizeMuncEBTrack(*args, **kw)
                np.testing.assert_array_equal(got[0], want[0])
                assert got[1] == want[1]
    out, diag = cb.cFinalizeMuncEBTrack(np.zeros(0, np.float32), np.zeros(0, np.float32), nuLocal=1.0, nuPrior=1.0)
    assert out.shape == (0,) and diag["supportFraction"] == 0.0 and diag["finalShrinkagePairFraction"] == 0.0


def test_finalize_errors_follow_the_reference(cb, oracle):
    rng = np.random.default_rng(4)
    n = 100_000
    loc = rng.uniform(1e-3, 2.0, n).astype(np.float32)
    pri = rng.uniform(1e-3, 2.0, n).astype(np.float32)
    cf = rng.uniform(0, 1, n).astype(np.float32)
    kw = dict(nuLocal=3.0, nuPrior=2.0, varianceFloor=1e-3, varianceCap=5.0)
    bad_l, bad_p, bad_c = loc.copy(), pri.copy(), cf.copy()
    bad_l[[90_000, 77_777]] = [np.nan, -1.0]
    bad_p[[300, 50_000]] = np.inf
    bad_c[[300, 301]] = [-1.0, np.inf]
    cases = [(bad_l, pri, cf), (loc, bad_p, cf), (loc, pri, bad_c), (bad_l, bad_p, bad_c), (loc, bad_p, bad_c)]
    for args in cases:
        msgs = []
        for mod in (oracle, cb):
            with pytest.raises(ValueError) as e:
                mod.cFinalizeMuncEBTrack(*args, **kw)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1], msgs
    for bad_kw, text in ((dict(varianceFloor=0.0), "varianceFloor must be positive and finite"),
                         (dict(varianceCap=1e-4), "varianceCap must be finite and at least varianceFloor"),
                         (dict(nuLocal=0.0), "nuLocal must be positive and finite"),
                         (dict(nuPrior=float("inf")), "nuPrior must be positive and finite")):
        with pytest.raises(ValueError, match=text):
            cb.cFinalizeMuncEBTrack(loc, pri, cf, **{**kw, **bad_kw})
    with pytest.raises(ValueError, match="priorVarianceTrack is required"):
        cb.cFinalizeMuncEBTrack(loc, None, cf, **kw)
    with pytest.raises(ValueError, match="priorVarianceTrack length must match"):
        cb.cFinalizeMuncEBTrack(loc, pri[:-1], cf, **kw)
    with pytest.raises(ValueError, match="countFloor length must match"):
        cb.cFinalizeMuncEBTrack(loc, pri, cf[:-1], **kw)


def test_concurrent_calls_from_a_thread_pool(cb, oracle):
    """The reference's MUNC stage calls these functions from a ThreadPool (consenrich.py:9055); every
    thread gets its own context, so concurrent calls of different shapes do not disturb one another."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(9)
    jobs = []
    for t in range(8):
        m, n, w = 2 + t % 3, 20_000 + 7_001 * t, 5 + 3 * t
        le = local_evidence(rng, m, n)
        jobs.append((le, w, oracle.cMuncSmoothDenseLocalEvidence(le, w)))

    def work(job):
        le, w, want = job
        for _ in range(5):
            assert_one_ulp(cb.cMuncSmoothDenseLocalEvidence(le, w), want, f"thread job window {w}")
        return True

    with ThreadPoolExecutor(max_workers=4) as pool:
        assert all(pool.map(work, jobs))


# ---- cMuncObservationMomentSeedPass: per-interval loops over the tracks, bit-identical ----
SEED_NAMES = ("moment", "rhoOut", "omegaRaw", "omegaOut", "local", "variance")


def test_seed_pass_matches_golden_vectors_bitwise(cb):
    from test_munc_oracle import check_seed, run_seed
    for name, c in golden_cases("seed").items():
        check_seed(run_seed(cb, c), c, name)


@pytest.mark.parametrize("variant", ["update", "fixed", "unweighted", "gaussian"])
def test_seed_pass_matches_oracle_bitwise_on_fresh_seeds(cb, oracle, variant):
    from golden.make_munc_golden import SEED_POSITIONAL, seed_case
    rng = np.random.default_rng(400 + len(variant))
    for m, n in ((1, 1), (2, 33), (10, 70001), (130, 4000), (3, 600_001)):
        c = seed_case(rng, m, n, variant)
        pos = [c[k] for k in SEED_POSITIONAL]
        kw = {k: v for k, v in c.items() if k not in SEED_POSITIONAL}
        want = oracle.cMuncObservationMomentSeedPass(*pos, **kw)
        got = cb.cMuncObservationMomentSeedPass(*pos, **kw)
        for name, g, w in zip(SEED_NAMES, got, want):
            assert g.shape == w.shape and g.dtype == np.float32
            np.testing.assert_array_equal(g, w, err_msg=f"{variant} {m}x{n} {name}")


def test_seed_pass_at_the_bench_size(cb, oracle):
    """hg38 chr19 @ 25 bp x 10 tracks: the oracle on the first 200 000 intervals (columns are independent),
    and the definitions themselves on the rest (variance = local + count floor inside the clip range)."""
    from golden.make_munc_golden import SEED_POSITIONAL, seed_case
    rng = np.random.default_rng(5)
    m, n, h = 10, 2_344_705, 200_000
    c = seed_case(rng, m, n, "update")
    pos = [c[k] for k in SEED_POSITIONAL]
    kw = {k: v for k, v in c.items() if k not in SEED_POSITIONAL}
    got = cb.cMuncObservationMomentSeedPass(*pos, **kw)
    head = lambda v: np.ascontiguousarray(v[..., :h]) if isinstance(v, np.ndarray) else v
    want = oracle.cMuncObservationMomentSeedPass(*[head(p) for p in pos], **{k: head(v) for k, v in kw.items()})
    for name, g, w in zip(SEED_NAMES, got, want):
        np.testing.assert_array_equal(g[..., :h], w, err_msg=name)
    local, variance = got[4].astype(np.float64), got[5].astype(np.float64)
    assert local.min() >= np.float32(1e-3) and variance.max() <= np.float32(1.5)
    inside = variance < 1.5 - 1e-6
    np.testing.assert_allclose(variance[inside], (local + c["countFloor"])[inside], rtol=2e-7)
    assert np.all((got[3] >= np.float32(0.5)) & (got[3] <= np.float32(1.5)))


def test_seed_pass_errors_follow_the_reference(cb, oracle):
    from golden.make_munc_golden import SEED_POSITIONAL, seed_case
    rng = np.random.default_rng(6)
    c = seed_case(rng, 3, 5000, "update")
    pos = [c[k] for k in SEED_POSITIONAL]
    kw = {k: v for k, v in c.items() if k not in SEED_POSITIONAL}
    k_act, k_off = int(np.flatnonzero(c["activeMask"])[-1]), int(np.flatnonzero(c["activeMask"] == 0)[0])
    bad_data, ok_data = pos[0].copy(), pos[0].copy()
    bad_data[1, k_act] = np.inf
    ok_data[1, k_off] = np.nan  # an inactive cell may hold anything
    bad_mean = pos[2].copy()
    bad_mean[k_act] = np.nan
    cases = [((bad_data, *pos[1:]), kw), ((pos[0], pos[1], bad_mean, pos[3]), kw), (pos, {**kw, "pad": -1.0}),
             (pos, {**kw, "varianceFloor": 0.0}), (pos, {**kw, "varianceCap": 1e-9}), (pos, {**kw, "omegaMin": 0.0}),
             (pos, {**kw, "omegaIn": c["omegaIn"][:-1]}), (pos, {**kw, "countFloor": c["countFloor"][:, :-1]}),
             (pos, {**kw, "activeMask": np.zeros((1, 2, 3), np.uint8)}), ((pos[0], pos[1][:, :-1], pos[2], pos[3]), kw)]
    for args, kwargs in cases:
        msgs = []
        for mod in (oracle, cb):
            with pytest.raises(ValueError) as e:
                mod.cMuncObservationMomentSeedPass(*args, **kwargs)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1], msgs
    a = oracle.cMuncObservationMomentSeedPass(ok_data, *pos[1:], **kw)
    b = cb.cMuncObservationMomentSeedPass(ok_data, *pos[1:], **kw)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


# ---- cEMA: affine scan + warm-up replay of the reference's own arithmetic ----
def assert_ema_close(got, want, what):
    """The reference's own tolerance for this function (tests/test_core.py:1364-1368): 1e-6 relative for
    float32, 1e-14 for float64 (scaled by the track's magnitude)."""
    assert got.dtype == want.dtype and got.shape == want.shape, what
    scale = max(float(np.abs(want).max()), 1e-30)
    if want.dtype == np.float32:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7 * scale, err_msg=what)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-14 * scale, err_msg=what)


def test_ema_matches_golden_vectors(cb):
    for name, c in golden_cases("ema").items():
        assert_ema_close(cb.cEMA(c["x"], float(c["alpha"])), c["out"], name)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ema_matches_oracle_on_fresh_seeds(cb, oracle, dtype):
    rng = np.random.default_rng(77)
    exact = total = 0
    for n in (1, 2, 255, 256, 257, 513, 70_001, 2_344_705):
        for alpha in (0.35, 2.0 / 42.0, 2.0 / 2001.0, 1.0, 0.0):
            x = (0.5 * np.sin(np.arange(n) / 50.0) + rng.normal(size=n)).astype(dtype)
            want, got = oracle.cEMA(x, alpha), cb.cEMA(x, alpha)
            assert_ema_close(got, want, f"n={n} alpha={alpha}")
            if alpha >= 2.0 / 42.0:  # (1 - alpha)^256 < 4e-6: the warm-up forgets the scan's start value
                exact += int(np.count_nonzero(got == want))
                total += n
    assert exact >= (1 - 1e-4) * total  # ... and then the replay is the reference's loop, bit for bit
    assert cb.cEMA(np.arange(5), 0.5).dtype == np.float64
    assert cb.cEMA(np.zeros(0, dtype), 0.5).shape == (0,)
    with pytest.raises(ValueError, match="alpha must lie in"):
        cb.cEMA(np.ones(4, dtype), 1.5)
