"""bedGraph writer (csrc/writer_kernels.cu) against the reference's own writer call -- pandas
``to_csv(sep="\\t", header=False, index=False, float_format="%.4f", lineterminator="\\n")``
(consenrich.py:9797-9805) -- byte for byte."""
import io
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def reference_chunk(chrom, starts, ends, values):
    df = pd.DataFrame({"Chromosome": [chrom] * len(values), "Start": starts, "End": ends, "v": values})
    buf = io.StringIO()
    df[["Chromosome", "Start", "End", "v"]].to_csv(buf, sep="\t", header=False, index=False, float_format="%.4f",
                                                   lineterminator="\n")
    return buf.getvalue().encode()


@pytest.fixture(scope="module")
def W():
    import consenrich_b200 as cb
    cb._lib.default_context(0)
    from consenrich_b200 import writers
    return writers


def edge_values():
    rng = np.random.default_rng(4)
    ties = (2 * np.arange(0, 4000, 7) + 1) * 625.0 / 20000.0 / 625.0  # not all dyadic; the dyadic ones are exact ties
    exact_ties = np.array([1 / 32, 3 / 32, 5 / 32, -1 / 32, 0.15625, 2.03125, 1000.09375], np.float64)
    specials = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-5, -1e-5, 4.9999e-5, 5.0001e-5, 0.99995, 0.999949,
                         9.99995, 16777215.0, 16777216.0, 16777218.0, 1e15, -1e15, 3.4028235e38, -3.4028235e38,
                         1.17549435e-38, 1e-45, 123456.789, -98765.4321], np.float64)
    rnd = np.concatenate([rng.normal(0, 3, 20000), rng.normal(0, 1e-4, 5000), 10.0 ** rng.uniform(-6, 12, 5000),
                          -(10.0 ** rng.uniform(-6, 30, 2000))])
    return np.concatenate([ties, exact_ties, specials, rnd]).astype(np.float32)


def test_bedgraph_chunk_is_byte_identical_to_pandas(W):
    v = edge_values()
    n = len(v)
    starts = np.arange(n, dtype=np.int64) * 25 + 10_000
    ends = starts + 25
    want = reference_chunk("chr7", starts, ends, v)
    assert W.bedgraph_chunk("chr7", v, starts, ends) == want
    assert W.bedgraph_chunk("chr7", v, start0=10_000, step=25) == want
    # the level column of a [n, 2] state array, ragged last interval, long names, large coordinates
    state = np.stack([v, v[::-1]], axis=1).copy()
    ends2 = ends.copy()
    ends2[-1] = ends[-1] - 7
    want2 = reference_chunk("chrUn_KI270442v1_random", starts + 2_400_000_000, ends2 + 2_400_000_000, v)
    got2 = W.bedgraph_chunk("chrUn_KI270442v1_random", state, start0=10_000 + 2_400_000_000, step=25,
                            end_clip=int(ends2[-1]) + 2_400_000_000)
    assert got2 == want2
    assert W.bedgraph_chunk("chr1", np.empty(0, np.float32), start0=0, step=25) == b""
    for k in (1, 255, 256, 257):  # tile boundaries
        assert W.bedgraph_chunk("chrX", v[:k], start0=0, step=50) == reference_chunk("chrX", np.arange(k) * 50, np.arange(k) * 50 + 50, v[:k])


def test_bedgraph_file_append_matches_reference_sequence(W, tmp_path):
    """Chromosome-ordered append: "w" for the first chromosome, "a" afterwards (consenrich.py:9802)."""
    rng = np.random.default_rng(0)
    ours, theirs = tmp_path / W.bedgraph_path("exp", "state", "0.0.0"), tmp_path / "ref.bedGraph"
    assert os.path.basename(ours) == "consenrichOutput_exp_state.v0.0.0.bedGraph"
    for c, (chrom, n) in enumerate((("chr1", 300_001), ("chr2", 70_000), ("chrM", 663))):
        v = rng.normal(0, 2, n).astype(np.float32)
        starts = np.arange(n, dtype=np.int64) * 25
        W.write_bedgraph_chunk(str(ours), chrom, v, first=(c == 0), start0=0, step=25)
        df = pd.DataFrame({"Chromosome": chrom, "Start": starts, "End": starts + 25, "v": v})
        df.to_csv(theirs, sep="\t", header=False, index=False, mode="w" if c == 0 else "a", float_format="%.4f",
                  lineterminator="\n")
    assert ours.read_bytes() == theirs.read_bytes()


def test_bedgraph_rejects_what_it_would_print_differently(W):
    with pytest.raises(TypeError):
        W.bedgraph_chunk("chr1", np.zeros(4, np.float64), start0=0, step=25)
    with pytest.raises(ValueError):
        W.bedgraph_chunk("chr1", np.zeros(4, np.float32), starts=np.arange(4))
    with pytest.raises(ValueError):
        W.bedgraph_chunk("c" * 40, np.zeros(4, np.float32), start0=0, step=25)
