"""bigWig writer (consenrich_b200/bigwig.py; the reference's io.py:633-780 + pyBigWig's container).

CPU only.  The checker is ``read_bigwig``, which finds everything through the file's own offsets, plus
structural checks written here against the published layout (magic numbers, header sizes, R-index covering
every section, zoom summaries agreeing with the data)."""
import os
import struct
import zlib

import numpy as np
import pytest

from consenrich_b200 import bigwig


SIZES = [("chr1", 1_000_000), ("chr2", 800_000), ("chrX", 500_000)]


def _track(rng, size, step, n, gaps=False):
    starts = np.arange(n, dtype=np.int64) * step
    if gaps:
        starts = starts[rng.random(n) > 0.3]
    ends = np.minimum(starts + step, size)
    keep = ends > starts
    starts, ends = starts[keep], ends[keep]
    values = np.round(rng.normal(size=len(starts)) * 3.0, 4).astype(np.float32)
    return starts, ends, values


def _write(tmp_path, tracks, sizes=SIZES, **kw):
    path = str(tmp_path / "out.bw")
    bigwig.write_bigwig(path, sizes, tracks, **kw)
    return path


def test_round_trip_and_header(tmp_path):
    rng = np.random.default_rng(0)
    tracks = [(c, *_track(rng, s, 25, s // 25, gaps=(c == "chr2"))) for c, s in SIZES]
    path = _write(tmp_path, tracks)
    raw = open(path, "rb").read()
    magic, version, n_zoom = struct.unpack_from("<IHH", raw, 0)
    assert magic == 0x888FFC26 and version == 4 and 1 <= n_zoom <= 10
    got = bigwig.read_bigwig(path)
    assert got["chroms"] == SIZES and got["end_signature"]
    assert (got["field_count"], got["defined_field_count"], got["autosql_offset"]) == (0, 0, 0)
    n_items = 0
    for c, s, e, v in tracks:
        gs, ge, gv = got["tracks"][c]
        np.testing.assert_array_equal(gs, s)
        np.testing.assert_array_equal(ge, e)
        np.testing.assert_array_equal(gv, v)  # float32 in, float32 stored: bit-exact
        n_items += len(s)
    assert got["sections"] == sum((len(t[1]) + 1023) // 1024 for t in tracks) == got["index"]["count"]
    assert got["index"]["items_per_slot"] == 1024
    width = np.concatenate([t[2] - t[1] for t in tracks]).astype(np.float64)
    vals = np.concatenate([t[3] for t in tracks]).astype(np.float64)
    summ = got["summary"]
    assert summ["bases_covered"] == int(width.sum())
    assert summ["min"] == vals.min() and summ["max"] == vals.max()
    np.testing.assert_allclose(summ["sum"], (vals * width).sum(), rtol=1e-12)
    np.testing.assert_allclose(summ["sum_squares"], (vals * vals * width).sum(), rtol=1e-12)


def test_index_is_multi_level_and_ordered(tmp_path):
    # > 256 sections forces an inner R-tree level; > 65536 would force a third (too slow for a unit test)
    rng = np.random.default_rng(1)
    sizes = [("chrA", 40_000_000), ("chrB", 10_000_000)]
    tracks = [(c, *_track(rng, s, 100, s // 100)) for c, s in sizes]
    path = _write(tmp_path, tracks, sizes=sizes, zoom_levels=2)
    got = bigwig.read_bigwig(path)  # the reader checks every child against its parent's bounding box
    leaves = got["index"]["leaves"]
    assert len(leaves) == 400_000 // 1024 + 1 + 100_000 // 1024 + 1 > 256
    raw = open(path, "rb").read()
    index_off = struct.unpack_from("<Q", raw, 24)[0]
    is_leaf = raw[index_off + 48]
    assert is_leaf == 0  # root is an inner node
    keys = [(a, b) for a, b, *_ in leaves]
    assert keys == sorted(keys)
    # sections are laid out back to back from the data offset, and the index says where the data ends
    data_off = struct.unpack_from("<Q", raw, 16)[0]
    pos = data_off + 8
    for _a, _b, _c, _d, off, size in leaves:
        assert off == pos
        pos += size
    assert pos == index_off == got["index"]["end_file"]
    assert got["index"]["bounds"] == (0, 0, 1, 10_000_000)
    for c, s, e, v in tracks:
        np.testing.assert_array_equal(got["tracks"][c][0], s)
        np.testing.assert_array_equal(got["tracks"][c][2], v)


def test_zoom_levels_summarise_the_data(tmp_path):
    rng = np.random.default_rng(2)
    tracks = [(c, *_track(rng, s, 25, s // 25, gaps=True)) for c, s in SIZES]
    path = _write(tmp_path, tracks)
    got = bigwig.read_bigwig(path)
    assert len(got["zooms"]) >= 3
    reductions = [z["reduction"] for z in got["zooms"]]
    assert reductions == sorted(reductions) and reductions[0] == 250 and all(b == 4 * a for a, b in zip(reductions, reductions[1:]))
    counts = [len(z["records"]) for z in got["zooms"]]
    assert all(b < a for a, b in zip(counts, counts[1:]))
    summ = got["summary"]
    for z in got["zooms"]:
        r = z["records"]
        assert int(r["valid"].astype(np.int64).sum()) == summ["bases_covered"]
        np.testing.assert_allclose(r["sum"].astype(np.float64).sum(), summ["sum"], rtol=1e-4, atol=1e-2 * np.sqrt(len(r)))
        np.testing.assert_allclose(r["sumsq"].astype(np.float64).sum(), summ["sum_squares"], rtol=1e-4)
        assert np.float32(r["min"].min()) == np.float32(summ["min"]) and np.float32(r["max"].max()) == np.float32(summ["max"])
        assert np.all(r["end"] - r["start"] == z["reduction"]) and np.all(r["start"] % z["reduction"] == 0)
        key = r["chrom"].astype(np.int64) * (1 << 32) + r["start"]
        assert np.all(np.diff(key) > 0)
    # one window recomputed by hand
    z = got["zooms"][1]
    rec = z["records"][len(z["records"]) // 2]
    name = got["chroms"][int(rec["chrom"])][0]
    s, e, v = got["tracks"][name]
    lo, hi = np.maximum(s, rec["start"]), np.minimum(e, rec["end"])
    cov = np.clip(hi - lo, 0, None)
    assert int(cov.sum()) == int(rec["valid"])
    np.testing.assert_allclose((v[cov > 0].astype(np.float64) * cov[cov > 0]).sum(), rec["sum"], rtol=1e-5, atol=1e-3)


def test_uncompressed_and_no_zoom(tmp_path):
    rng = np.random.default_rng(3)
    tracks = [("chr2", *_track(rng, 800_000, 50, 3000))]
    path = _write(tmp_path, tracks, zoom_levels=0, compress=False)
    got = bigwig.read_bigwig(path)
    assert got["uncompress_buf_size"] == 0 and got["zooms"] == []
    np.testing.assert_array_equal(got["tracks"]["chr2"][2], tracks[0][3])
    assert len(got["tracks"]["chr1"][0]) == 0
    # compressed sections inflate to at most the header's buffer size
    path2 = str(tmp_path / "c.bw")
    bigwig.write_bigwig(path2, SIZES, tracks)
    raw = open(path2, "rb").read()
    buf_size = struct.unpack_from("<I", raw, 52)[0]
    assert buf_size == 24 + 12 * 1024
    for _a, _b, _c, _d, off, size in bigwig.read_bigwig(path2)["index"]["leaves"]:
        assert len(zlib.decompress(raw[off:off + size])) <= buf_size


def test_chromosome_tree_keys_sorted(tmp_path):
    sizes = [("chr10", 5000), ("chr2", 7000), ("chr1", 9000), ("chrUn_KI270742v1", 3000)]
    tracks = [("chr2", np.array([0, 100]), np.array([100, 200]), np.array([1.5, -2.0], np.float32)),
              ("chrUn_KI270742v1", np.array([2900]), np.array([3000]), np.array([7.0], np.float32))]
    path = _write(tmp_path, tracks, sizes=sizes)
    raw = open(path, "rb").read()
    chrom_off = struct.unpack_from("<Q", raw, 8)[0]
    magic, block, key_size, val_size, count, _ = struct.unpack_from("<IIIIQQ", raw, chrom_off)
    assert (magic, key_size, val_size, count) == (0x78CA8C91, len("chrUn_KI270742v1"), 8, 4)
    pos = chrom_off + 32 + 4
    keys = []
    for _ in range(count):
        keys.append(raw[pos:pos + key_size])
        pos += key_size + 8
    assert keys == sorted(keys)
    got = bigwig.read_bigwig(path)
    assert got["chroms"] == sizes  # ids follow the chromosome-sizes order
    np.testing.assert_array_equal(got["tracks"]["chrUn_KI270742v1"][1], [3000])


def _bedgraph(tmp_path, text, name="in.bedGraph"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_convert_bedgraph(tmp_path):
    sizes_path = tmp_path / "g.sizes"
    sizes_path.write_text("# comment\nchr1\t1000\nchr2 500 extra\n\n")
    assert bigwig.read_chrom_sizes(str(sizes_path)) == [("chr1", 1000), ("chr2", 500)]
    bg = _bedgraph(tmp_path, "track type=bedGraph\n#c\nchr1\t0\t25\t0.1250\nchr1\t25\t50\t-3.0000\n\nchr2\t100\t125\t1e-3\n")
    out = str(tmp_path / "x.bw")
    bigwig.convert_bedgraph_to_bigwig(bg, str(sizes_path), out)
    got = bigwig.read_bigwig(out)
    np.testing.assert_array_equal(got["tracks"]["chr1"][0], [0, 25])
    np.testing.assert_array_equal(got["tracks"]["chr1"][2], np.array([0.125, -3.0], np.float32))
    np.testing.assert_array_equal(got["tracks"]["chr2"][2], np.array([1e-3], np.float32))
    assert [f for f in os.listdir(tmp_path) if f.startswith("consenrich_bigwig_")] == []  # temp file moved into place
    assert bigwig.bigwig_path("exp", "state", "0.9.1") == "exp_consenrich_state.v0.9.1.bw"


@pytest.mark.parametrize("text, message", [
    ("chr1\t0\t25\n", "expected 4 columns"),
    ("chr9\t0\t25\t1\n", "Chromosome chr9 on bedGraph row 1 is not present"),
    ("chr1\ta\t25\t1\n", "Invalid bedGraph coordinates on row 1"),
    ("chr1\t0\t25\tx\n", "Invalid bedGraph value on row 1"),
    ("chr1\t0\t25\tnan\n", "Non-finite bedGraph value on row 1"),
    ("chr1\t-5\t25\t1\n", "Negative start coordinate on bedGraph row 1"),
    ("chr1\t25\t25\t1\n", "End coordinate must be greater than start on bedGraph row 1"),
    ("chr1\t0\t2000\t1\n", "End coordinate 2000 on bedGraph row 1 exceeds chr1 size of 1000"),
    ("chr2\t0\t25\t1\nchr1\t0\t25\t1\n", "not sorted at row 2"),
    ("chr1\t50\t75\t1\nchr1\t25\t50\t1\n", "not sorted at row 2"),
    ("chr1\t0\t50\t1\nchr1\t25\t75\t1\n", "Overlapping bedGraph interval at row 2"),
    ("# nothing\n", "No bedGraph intervals found"),
])
def test_convert_rejects_what_the_reference_rejects(tmp_path, text, message):
    bg = _bedgraph(tmp_path, text)
    out = str(tmp_path / "bad.bw")
    with pytest.raises(ValueError, match=message):
        bigwig.convert_bedgraph_to_bigwig(bg, [("chr1", 1000), ("chr2", 500)], out)
    assert not os.path.exists(out)
    assert [f for f in os.listdir(tmp_path) if f.startswith("consenrich_bigwig_")] == []


def test_write_rejects_bad_tracks(tmp_path):
    ok = (np.array([0, 25]), np.array([25, 50]), np.array([1.0, 2.0], np.float32))
    with pytest.raises(ValueError, match="order"):
        _write(tmp_path, [("chr2", *ok), ("chr1", *ok)])
    with pytest.raises(ValueError, match="not present"):
        _write(tmp_path, [("chrZ", *ok)])
    with pytest.raises(ValueError, match="Overlapping"):
        _write(tmp_path, [("chr1", np.array([0, 10]), np.array([25, 50]), ok[2])])
    with pytest.raises(ValueError, match="No intervals"):
        _write(tmp_path, [])


def test_chrom_sizes_errors(tmp_path):
    for text, message in [("chr1\n", "Malformed"), ("chr1 x\n", "Invalid chromosome size"), ("chr1 0\n", "non-positive"),
                          ("chr1 5\nchr1 6\n", "Duplicate"), ("#\n", "No chromosome sizes")]:
        p = tmp_path / "s.sizes"
        p.write_text(text)
        with pytest.raises(ValueError, match=message):
            bigwig.read_chrom_sizes(str(p))


def test_convert_outputs_loop(tmp_path):
    sizes = tmp_path / "g.sizes"
    sizes.write_text("chr1 1000\nchr2 500\n")
    d = str(tmp_path)
    (tmp_path / "consenrichOutput_exp_state.v1.2.3.bedGraph").write_text("chr1\t0\t25\t1.0000\nchr2\t0\t25\t2.0000\n")
    unsorted = tmp_path / "consenrichOutput_exp_uncertainty.v1.2.3.bedGraph"
    unsorted.write_text("track type=bedGraph\nchr2\t0\t25\t1\nchr1\t50\t75\t0.12345\nchr1\t0\t25\t-3\n")
    (tmp_path / "consenrichOutput_exp_bad.v1.2.3.bedGraph").write_text("chr1\t0\t50\t1\nchr1\t25\t75\t1\n")  # overlap
    with pytest.warns(UserWarning) as rec:
        written = bigwig.convert_outputs("exp", str(sizes), ["state", "uncertainty", "bad", "missing"], version="1.2.3",
                                         delete_bedgraphs=True, directory=d)
    messages = " | ".join(str(w.message) for w in rec)
    assert "sorting as a fallback" in messages and "not sorted at row 3" in messages
    assert "Overlapping bedGraph interval at row 2" in messages and "missing.v1.2.3.bedGraph does not exist" in messages
    assert written == [os.path.join(d, "exp_consenrich_state.v1.2.3.bw"), os.path.join(d, "exp_consenrich_uncertainty.v1.2.3.bw")]
    assert not os.path.exists(tmp_path / "consenrichOutput_exp_state.v1.2.3.bedGraph")   # converted and deleted
    assert os.path.exists(tmp_path / "consenrichOutput_exp_bad.v1.2.3.bedGraph")          # failed: kept
    np.testing.assert_array_equal(bigwig.read_bigwig(written[0])["tracks"]["chr2"][2], np.array([2.0], np.float32))
    got = bigwig.read_bigwig(written[1])["tracks"]  # the unsorted file was repaired the way the reference repairs it
    np.testing.assert_array_equal(got["chr1"][0], [0, 50])
    np.testing.assert_array_equal(got["chr1"][2], np.array([-3.0, 0.1235], np.float32))  # values re-printed as %.4f
    with pytest.warns(UserWarning, match="does not exist"):
        assert bigwig.convert_outputs("exp", str(tmp_path / "nope.sizes"), ["bad"], version="1.2.3", directory=d) == []


def test_sort_bedgraph_in_place(tmp_path):
    p = _bedgraph(tmp_path, "# c\nchr2\t10\t20\t1\nchr1\t30\t40\t2.00006\nchr1\t30\t35\t-0.00004\ntrack x\nchr1\t0\t5\t1e2\n")
    bigwig.sort_bedgraph_in_place(p, ["chr1", "chr2"])
    assert open(p).read() == ("# c\ntrack x\nchr1\t0\t5\t100.0000\nchr1\t30\t35\t-0.0000\nchr1\t30\t40\t2.0001\n"
                              "chr2\t10\t20\t1.0000\n")
    with pytest.raises(ValueError, match="not present in chromosome order: chr2"):
        bigwig.sort_bedgraph_in_place(p, ["chr1"])
    assert [f for f in os.listdir(tmp_path) if f.startswith("consenrich_sort_")] == []
    # a validated file is never rewritten: the failure is reported instead
    q = tmp_path / "consenrichOutput_e_state.v1.bedGraph"
    q.write_text("chr2\t0\t25\t1\nchr1\t0\t25\t1\n")
    before = q.read_text()
    (tmp_path / "s.sizes").write_text("chr1 100\nchr2 100\n")
    with pytest.warns(UserWarning, match="not sorted at row 2"):
        assert bigwig.convert_outputs("e", str(tmp_path / "s.sizes"), version="1", directory=str(tmp_path), validated=[str(q)]) == []
    assert q.read_text() == before


def test_values_as_printed_equal_the_text_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    v = np.concatenate([rng.normal(size=20000) * 10.0 ** rng.integers(-6, 5, size=20000),
                        (rng.integers(-10 ** 6, 10 ** 6, size=5000) * 2 + 1) / 2.0e4,  # decimal ties of the 4th digit
                        [0.0, -0.0, 1e-5, -4.99999e-5, 5e-5, 16777216.0, 123456.789]]).astype(np.float32)
    want = np.array([np.float32(float("%.4f" % x)) for x in v.tolist()], dtype=np.float32)
    np.testing.assert_array_equal(bigwig.values_as_printed(v), want)
    # and so the two ways into a bigWig give the same intervals and values
    vals = v[:3000]
    chrom, s, e, q = bigwig.fixed_step_track("chr1", vals, start0=100, step=25, chrom_size=75_090)
    assert e[-1] == 75_090 and s[0] == 100 and e[-2] - s[-2] == 25
    direct = str(tmp_path / "direct.bw")
    bigwig.write_bigwig(direct, SIZES, [(chrom, s, e, q)])
    text = "".join(f"chr1\t{a}\t{b}\t{x:.4f}\n" for a, b, x in zip(s.tolist(), e.tolist(), vals.tolist()))
    via_text = str(tmp_path / "text.bw")
    bigwig.convert_bedgraph_to_bigwig(_bedgraph(tmp_path, text), SIZES, via_text)
    assert open(direct, "rb").read() == open(via_text, "rb").read()


def test_chromosome_tree_with_many_contigs(tmp_path):
    # 70 000 contigs: more than one node's 16-bit count could hold, three levels of 256-way nodes
    rng = np.random.default_rng(11)
    names = [f"ctg{i:06d}_{rng.integers(1 << 30):x}" for i in range(70_000)]
    order = rng.permutation(len(names))
    sizes = [(names[i], 1000 + int(i)) for i in order]
    pick = [sizes[0], sizes[12345], sizes[-1]]
    tracks = [(c, np.array([0]), np.array([s]), np.array([float(s)], np.float32)) for c, s in pick]
    path = _write(tmp_path, tracks, sizes=sizes, zoom_levels=0)
    raw = open(path, "rb").read()
    chrom_off = struct.unpack_from("<Q", raw, 8)[0]
    magic, block, key_size, val_size, count, _ = struct.unpack_from("<IIIIQQ", raw, chrom_off)
    assert (magic, block, val_size, count) == (0x78CA8C91, 256, 8, 70_000)
    assert raw[chrom_off + 32] == 0  # the root is an inner node

    def lookup(name):  # the search a reader does: last key <= name on every inner level
        key = name.encode().ljust(key_size, b"\0")
        pos = chrom_off + 32
        while True:
            is_leaf, _r, n = struct.unpack_from("<BBH", raw, pos)
            pos += 4
            entries = [(raw[pos + i * (key_size + 8):pos + i * (key_size + 8) + key_size],
                        raw[pos + i * (key_size + 8) + key_size:pos + (i + 1) * (key_size + 8)]) for i in range(n)]
            assert [k for k, _ in entries] == sorted(k for k, _ in entries)
            if is_leaf:
                hit = [v for k, v in entries if k == key]
                return struct.unpack("<II", hit[0]) if hit else None
            below = [v for k, v in entries if k <= key]
            if not below:
                return None
            pos = struct.unpack("<Q", below[-1])[0]
    for probe in (0, 1, 255, 256, 65_535, 65_536, 69_999):
        name, size = sizes[probe]
        assert lookup(name) == (probe, size)
    assert lookup("zzz") is None and lookup("a") is None
    got = bigwig.read_bigwig(path)
    assert got["chroms"] == sizes
    for c, s in pick:
        np.testing.assert_array_equal(got["tracks"][c][1], [s])


def test_region_queries_descend_the_index(tmp_path):
    rng = np.random.default_rng(5)
    sizes = [("chrA", 40_000_000), ("chrB", 10_000_000), ("chrC", 1_000_000)]
    tracks = [(c, *_track(rng, s, 100, s // 100, gaps=True)) for c, s in sizes]
    path = _write(tmp_path, tracks, sizes=sizes, zoom_levels=1)
    by_name = {c: (s, e, v) for c, s, e, v in tracks}
    for _ in range(60):
        c, size = sizes[int(rng.integers(3))]
        a = int(rng.integers(0, size - 1))
        b = min(size, a + int(10 ** rng.uniform(0, 6)))
        s, e, v = by_name[c]
        keep = (s < b) & (e > a)
        gs, ge, gv = bigwig.query_bigwig(path, c, a, b)
        np.testing.assert_array_equal(gs, s[keep])
        np.testing.assert_array_equal(ge, e[keep])
        np.testing.assert_array_equal(gv, v[keep])
    assert len(bigwig.query_bigwig(path, "chrC", 0, 1_000_000)[0]) == len(by_name["chrC"][0])
    with pytest.raises(KeyError):
        bigwig.query_bigwig(path, "chrD", 0, 10)
