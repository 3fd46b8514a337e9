"""bench.py's workload definitions (no GPU): the configurations are BASELINE.json's, the chromosome tables give
SURVEY 8's interval counts, both arms emit the same `config`, and the per-kernel algorithmic bytes add up."""
import json
import os
import subprocess
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_interval_counts_match_the_survey():
    tot = {k: sum(bench.chrom_bins(c).values()) for k, c in bench.CONFIGS.items()}
    assert tot == {"cfg2": 2_344_705, "cfg3": 123_530_804, "cfg4": 308_826_993, "cfg5": 61_765_409}
    assert bench.chrom_bins(bench.CONFIGS["cfg3"])["chr1"] == 9_958_257
    assert bench.chrom_bins(bench.CONFIGS["cfg4"])["chr1"] == 24_895_643
    assert [bench.CONFIGS[k]["m"] for k in ("cfg2", "cfg3", "cfg4", "cfg5")] == [10, 50, 200, 1000]
    assert len(bench.HG38) == 24


def test_config_object_is_the_same_for_both_arms_and_json_clean():
    for name in bench.CONFIGS:
        c = bench.make_config(name)
        assert json.loads(json.dumps(c)) == c
        assert c["m"] == bench.CONFIGS[name]["m"] and c["sweeps_per_step"] == bench.ECM_ITERS * bench.T_INNER
        assert name in c["workload"] and "l2" in c


def test_algorithmic_bytes_per_kernel():
    m, bins = 50, [9_958_257, 2_344_705]
    alg = bench.algorithmic_bytes(m, bins, True, lambda n: 5)
    n_all, K, t = float(sum(bins)), bench.ECM_ITERS, bench.T_INNER
    assert alg["fold"] == (m * n_all * 8 + n_all * 32, 2)
    assert alg["forward_scan"][1] == (K * t + 1) * 2 and alg["backward_scan"][1] == K * t * 2
    # a plain forward replay moves 52 B per interval plus the run elements; the whole step is dominated by
    # the sweeps, not by the one pass over the [m x n] matrices
    per_launch = alg["forward_scan"][0] / alg["forward_scan"][1] / n_all * 2
    assert 52.0 < per_launch < 70.0
    old = bench.algorithmic_bytes(m, bins, False, lambda n: 5)
    assert set(old) == {"fold", "residuals", "forward_scan", "backward_scan"}


def test_reference_arm_prints_one_json_line_with_the_shared_config(tmp_path):
    """The CPU arm end to end on the smallest configuration (a few seconds: one step of a bounded sample)."""
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "consenrich_ref")):
        import pytest
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                         env={**os.environ, "CB200_REF_BUDGET_S": "4"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["config"] == bench.make_config("cfg2") and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
