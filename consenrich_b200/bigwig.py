"""bigWig output: what ``consenrich.io.convertBedGraphToBigWig`` produces from the finished
bedGraph files (io.py:530-780).

The reference does not write the file format itself: it parses and validates the bedGraph
(``_convertBedGraphToBigWigPyBigWig``, io.py:633-780 -- four columns, known chromosome, 0 <= start < end <=
chromosome size, finite value, sorted by the chromosome-sizes order then start, no overlaps) and hands the rows
to the third-party **pyBigWig** (libBigWig; not vendored in the reference tree and not installed here), in
chunks of 200 000 via ``addHeader`` / ``addEntries``.  This module restates both halves on the host -- the
validation with the reference's messages, and the published BBI / bigWig container that libBigWig emits for
``addEntries(chroms, starts, ends=..., values=...)``:

    header (64 B) | zoom headers | total summary (40 B) | chromosome B+ tree | data count |
    zlib-compressed sections of <= 1024 bedGraph items (24 B section header + 12 B per item) |
    cirTree R-index over the sections | zoom levels (summary records + their own R-index) | end signature

(Kent et al., "BigWig and BigBed: enabling browsing of large distributed datasets", Bioinformatics 2010, and
the UCSC bbiFile.h / bwgInternal.h layouts).  It is host-side byte packing (numpy + zlib), not device work.
``read_bigwig`` is an independent reader of the same container: the tests round-trip through it, check the
R-index against the sections and the zoom summaries against the data.  Byte identity with libBigWig's own
output is not claimed (it cannot be checked here: pyBigWig is absent); interval / value identity is.
"""
from __future__ import annotations

import os
import struct
import tempfile
import zlib
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["write_bigwig", "convert_bedgraph_to_bigwig", "convert_outputs", "read_bigwig", "query_bigwig", "bigwig_path", "read_chrom_sizes",
           "values_as_printed", "fixed_step_track", "sort_bedgraph_in_place"]

BIGWIG_MAGIC = 0x888FFC26
BPT_MAGIC = 0x78CA8C91
CIRTREE_MAGIC = 0x2468ACE0
ITEMS_PER_SLOT = 1024      # bedGraph items per data section (libBigWig / UCSC default)
RTREE_BLOCK = 256          # children per R-tree node
ZOOM_RECORDS_PER_SLOT = 512
MAX_ZOOM_LEVELS = 10


def bigwig_path(experiment_name: str, suffix: str, version: str) -> str:
    """{experimentName}_consenrich_{suffix}.v{version}.bw (io.py:581)."""
    return f"{experiment_name}_consenrich_{suffix}.v{version}.bw"


def read_chrom_sizes(path: str) -> List[Tuple[str, int]]:
    """Rows ``name size`` in file order; the checks and messages of io.py:601-630."""
    out: List[Tuple[str, int]] = []
    seen: Dict[str, int] = {}
    with open(path, "r", encoding="utf-8") as handle:
        for line_number, line in enumerate(handle, start=1):
            parts = line.rstrip("\n").split()
            if len(parts) == 0 or parts[0].startswith("#"):
                continue
            if len(parts) < 2:
                raise ValueError(f"Malformed chromosome sizes row {line_number} in {path}")
            chrom = str(parts[0])
            try:
                size = int(parts[1])
            except ValueError as e:
                raise ValueError(f"Invalid chromosome size on row {line_number} in {path}") from e
            if size <= 0:
                raise ValueError(f"Chromosome {chrom} has non-positive size on row {line_number}")
            if chrom in seen:
                raise ValueError(f"Duplicate chromosome {chrom} in {path}")
            out.append((chrom, size))
            seen[chrom] = size
    if len(out) == 0:
        raise ValueError(f"No chromosome sizes found in {path}")
    return out


# ------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------
def _rtree_bytes(items: np.ndarray, file_offset: int, items_per_slot: int) -> bytes:
    """cirTree over ``items`` (structured: chrom, start, end_chrom, end, offset, size), written at
    ``file_offset``: header (48 B), then the levels from the root down, leaves last."""
    n = len(items)
    levels = [items]  # level 0 = leaf items; upper levels hold bounding boxes of groups of RTREE_BLOCK children
    while len(levels[-1]) > RTREE_BLOCK:
        cur = levels[-1]
        groups = (len(cur) + RTREE_BLOCK - 1) // RTREE_BLOCK
        up = np.zeros(groups, dtype=cur.dtype)
        for g in range(groups):
            sl = cur[g * RTREE_BLOCK:(g + 1) * RTREE_BLOCK]
            first, last = sl[0], sl[-1]
            up[g]["chrom"], up[g]["start"] = first["chrom"], first["start"]
            # bounding end: the largest (end_chrom, end) of the children
            k = np.lexsort((sl["end"], sl["end_chrom"]))[-1]
            up[g]["end_chrom"], up[g]["end"] = sl[k]["end_chrom"], sl[k]["end"]
        levels.append(up)
    # byte size of every level: nodes of <= RTREE_BLOCK children; leaf items 32 B, inner items 24 B
    def level_size(count, leaf):
        nodes = (count + RTREE_BLOCK - 1) // RTREE_BLOCK if count else 1
        return nodes * 4 + count * (32 if leaf else 24)
    header = 48
    # layout: root level first ... leaf level last
    order = list(range(len(levels) - 1, -1, -1))
    offsets = {}
    pos = file_offset + header
    for lv in order:
        offsets[lv] = pos
        pos += level_size(len(levels[lv]), lv == 0)
    out = bytearray()
    end_file = int(items[-1]["offset"] + items[-1]["size"]) if n else file_offset
    k_end = np.lexsort((items["end"], items["end_chrom"]))[-1] if n else 0
    out += struct.pack("<IIQIIIIQII", CIRTREE_MAGIC, RTREE_BLOCK, n,
                       int(items[0]["chrom"]) if n else 0, int(items[0]["start"]) if n else 0,
                       int(items[k_end]["end_chrom"]) if n else 0, int(items[k_end]["end"]) if n else 0,
                       end_file, items_per_slot, 0)
    for lv in order:
        cur = levels[lv]
        leaf = lv == 0
        count = len(cur)
        nodes = (count + RTREE_BLOCK - 1) // RTREE_BLOCK if count else 1
        # where node j of the level below starts
        if not leaf:
            below_leaf = lv - 1 == 0
            child_item = 32 if below_leaf else 24
            child_base = offsets[lv - 1]
        for j in range(nodes):
            sl = cur[j * RTREE_BLOCK:(j + 1) * RTREE_BLOCK]
            out += struct.pack("<BBH", 1 if leaf else 0, 0, len(sl))
            if leaf:
                rec = np.zeros(len(sl), dtype=[("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4"), ("o", "<u8"), ("s", "<u8")])
                rec["a"], rec["b"], rec["c"], rec["d"] = sl["chrom"], sl["start"], sl["end_chrom"], sl["end"]
                rec["o"], rec["s"] = sl["offset"], sl["size"]
                out += rec.tobytes()
            else:
                rec = np.zeros(len(sl), dtype=[("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4"), ("o", "<u8")])
                rec["a"], rec["b"], rec["c"], rec["d"] = sl["chrom"], sl["start"], sl["end_chrom"], sl["end"]
                # child g of this level is node (j * RTREE_BLOCK + g) of the level below
                g = np.arange(j * RTREE_BLOCK, j * RTREE_BLOCK + len(sl), dtype=np.int64)
                full = RTREE_BLOCK
                rec["o"] = child_base + g * (4 + full * child_item)
                out += rec.tobytes()
    assert len(out) == pos - file_offset, (len(out), pos - file_offset)
    return bytes(out)


_ITEM_DT = np.dtype([("chrom", "<u4"), ("start", "<u4"), ("end_chrom", "<u4"), ("end", "<u4"), ("offset", "<u8"), ("size", "<u8")])


_ZOOM_WORK_DT = np.dtype([("chrom", "<u4"), ("start", "<u4"), ("end", "<u4"), ("valid", "<u4"), ("min", "<f4"), ("max", "<f4"),
                          ("sum", "<f8"), ("sumsq", "<f8")])  # sums in float64 until the records are written


def _zoom_records(chrom_id: int, starts: np.ndarray, ends: np.ndarray, values: np.ndarray, reduction: int):
    """Summary records of one chromosome at the first zoom level: windows of ``reduction`` bases (aligned to
    multiples of it), each with the bases covered, min, max, sum and sum of squares of the data in it."""
    if len(starts) == 0:
        return np.zeros(0, dtype=_ZOOM_WORK_DT)
    v = values.astype(np.float64)
    first_w, last_w = starts // reduction, (ends - 1) // reduction
    span = int(np.max(last_w - first_w)) + 1
    # an interval contributes to every window it overlaps, weighted by the overlap (intervals are short
    # against any zoom window, so `span` is 1 or 2 in practice)
    parts_w, parts_cov, parts_v = [], [], []
    for s_ in range(span):
        w = first_w + s_
        ok = w <= last_w
        lo = np.maximum(starts, w * reduction)
        hi = np.minimum(ends, (w + 1) * reduction)
        cov = np.where(ok, hi - lo, 0)
        keep = cov > 0
        parts_w.append(w[keep])
        parts_cov.append(cov[keep])
        parts_v.append(v[keep])
    w = np.concatenate(parts_w)
    cov = np.concatenate(parts_cov).astype(np.float64)
    vv = np.concatenate(parts_v)
    if span > 1:
        order = np.argsort(w, kind="stable")
        w, cov, vv = w[order], cov[order], vv[order]
    idx = np.flatnonzero(np.concatenate(([True], w[1:] != w[:-1])))
    uniq = w[idx]
    out = np.zeros(len(uniq), dtype=_ZOOM_WORK_DT)
    out["chrom"] = chrom_id
    out["start"] = uniq * reduction
    out["end"] = (uniq + 1) * reduction
    out["valid"] = np.add.reduceat(cov, idx).astype(np.uint32)
    out["min"] = np.minimum.reduceat(vv, idx)
    out["max"] = np.maximum.reduceat(vv, idx)
    out["sum"] = np.add.reduceat(vv * cov, idx)
    out["sumsq"] = np.add.reduceat(vv * vv * cov, idx)
    return out


def _zoom_coarsen(recs: np.ndarray, reduction: int) -> np.ndarray:
    """The next zoom level from the records of the one below: windows nest (reductions grow by whole factors and
    are aligned to multiples of themselves), so a coarse window is the merge of the fine ones inside it."""
    if len(recs) == 0:
        return recs
    key = recs["chrom"].astype(np.int64) * (1 << 32) + recs["start"] // reduction
    idx = np.flatnonzero(np.concatenate(([True], key[1:] != key[:-1])))
    out = np.zeros(len(idx), dtype=_ZOOM_WORK_DT)
    out["chrom"] = recs["chrom"][idx]
    out["start"] = (recs["start"][idx] // reduction) * reduction
    out["end"] = out["start"] + reduction
    out["valid"] = np.add.reduceat(recs["valid"].astype(np.int64), idx).astype(np.uint32)
    out["min"] = np.minimum.reduceat(recs["min"], idx)
    out["max"] = np.maximum.reduceat(recs["max"], idx)
    out["sum"] = np.add.reduceat(recs["sum"], idx)
    out["sumsq"] = np.add.reduceat(recs["sumsq"], idx)
    return out


_ZOOM_DT = np.dtype([("chrom", "<u4"), ("start", "<u4"), ("end", "<u4"), ("valid", "<u4"), ("min", "<f4"), ("max", "<f4"),
                     ("sum", "<f4"), ("sumsq", "<f4")])


BPT_BLOCK = 256  # children per node of the chromosome B+ tree


def _chrom_tree_bytes(chrom_sizes, ids, key_size: int, file_offset: int) -> bytes:
    """Chromosome B+ tree (name -> id, size), written at ``file_offset``: header (32 B), then the levels from the
    root down; keys in byte order, an inner item is the first key of its child and the child's file offset."""
    items = sorted(((c.encode().ljust(key_size, b"\0"), ids[c], size) for c, size in chrom_sizes), key=lambda t: t[0])
    n = len(items)
    block = min(BPT_BLOCK, max(n, 1))
    counts = [n]  # items per level, leaves first
    while counts[-1] > block:
        counts.append((counts[-1] + block - 1) // block)
    item_bytes = key_size + 8

    def level_size(count):
        return ((count + block - 1) // block) * 4 + count * item_bytes
    offsets = {}
    pos = file_offset + 32
    for lv in range(len(counts) - 1, -1, -1):
        offsets[lv] = pos
        pos += level_size(counts[lv])
    out = bytearray(struct.pack("<IIIIQQ", BPT_MAGIC, block, key_size, 8, n, 0))
    for lv in range(len(counts) - 1, -1, -1):
        stride = block ** lv  # leaf items under one item of this level
        for j in range((counts[lv] + block - 1) // block):
            lo, hi = j * block, min((j + 1) * block, counts[lv])
            out += struct.pack("<BBH", 1 if lv == 0 else 0, 0, hi - lo)
            for g in range(lo, hi):
                if lv == 0:
                    key, cid, size = items[g]
                    out += key + struct.pack("<II", cid, size)
                else:
                    child = offsets[lv - 1] + g * (4 + block * item_bytes)  # node g of the level below (all full before it)
                    out += items[g * stride][0] + struct.pack("<Q", child)
    assert len(out) == pos - file_offset
    return bytes(out)


def _deflate_all(blocks: List[bytes]) -> List[bytes]:
    """zlib-compress every block; the calls release the GIL, so a few threads share them."""
    if len(blocks) < 64:
        return [zlib.compress(b) for b in blocks]
    from concurrent.futures import ThreadPoolExecutor
    workers = max(1, min(8, (os.cpu_count() or 1)))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(zlib.compress, blocks, chunksize=64))


def write_bigwig(path: str, chrom_sizes: Sequence[Tuple[str, int]],
                 tracks: Iterable[Tuple[str, np.ndarray, np.ndarray, np.ndarray]], *, zoom_levels: int = MAX_ZOOM_LEVELS,
                 compress: bool = True) -> None:
    """Write ``tracks`` -- (chromosome, starts, ends, values) per chromosome, in the order of ``chrom_sizes``,
    each sorted by start and free of overlaps -- as a bigWig of bedGraph-type sections.  The file is written
    to a temporary name in the target directory and moved into place (io.py:661-770 does the same)."""
    chrom_sizes = [(str(c), int(s)) for c, s in chrom_sizes]
    if len(chrom_sizes) == 0:
        raise ValueError("No chromosome sizes given")
    ids = {c: i for i, (c, _s) in enumerate(chrom_sizes)}
    size_of = dict(chrom_sizes)
    sections: List[bytes] = []
    sec_items: List[Tuple[int, int, int, int]] = []  # chrom, start, end, raw size
    per_chrom = []
    covered, vmin, vmax, vsum, vsumsq = 0, np.inf, -np.inf, 0.0, 0.0
    last_rank = -1
    max_raw = 0
    for chrom, starts, ends, values in tracks:
        if chrom not in ids:
            raise ValueError(f"Chromosome {chrom} is not present in the chromosome sizes")
        rank = ids[chrom]
        if rank <= last_rank:
            raise ValueError("tracks must follow the chromosome-sizes order, one entry per chromosome")
        last_rank = rank
        s = np.ascontiguousarray(starts, dtype=np.int64).reshape(-1)
        e = np.ascontiguousarray(ends, dtype=np.int64).reshape(-1)
        v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        if not (len(s) == len(e) == len(v)):
            raise ValueError("starts, ends and values must have one length")
        if len(s) == 0:
            continue
        if not np.all(np.isfinite(v)):
            raise ValueError(f"Non-finite value on {chrom}")
        if s[0] < 0 or np.any(e <= s) or e[-1] > size_of[chrom] or np.any(e > size_of[chrom]):
            raise ValueError(f"Interval outside 0 <= start < end <= {size_of[chrom]} on {chrom}")
        if np.any(s[1:] < s[:-1]):
            raise ValueError(f"Intervals of {chrom} are not sorted by start")
        if np.any(s[1:] < e[:-1]):
            raise ValueError(f"Overlapping intervals on {chrom}")
        per_chrom.append((rank, s, e, v))
        width = (e - s).astype(np.float64)
        v64 = v.astype(np.float64)
        covered += int(width.sum())
        vmin, vmax = min(vmin, float(v.min())), max(vmax, float(v.max()))
        vsum += float((v64 * width).sum())
        vsumsq += float((v64 * v64 * width).sum())
        # whole chromosome packed at once: [start, end, value] triples, cut into sections of ITEMS_PER_SLOT
        rec = np.zeros(len(s), dtype=[("s", "<u4"), ("e", "<u4"), ("v", "<f4")])
        rec["s"], rec["e"], rec["v"] = s, e, v
        body = rec.tobytes()
        raws = []
        for a in range(0, len(s), ITEMS_PER_SLOT):
            b = min(a + ITEMS_PER_SLOT, len(s))
            raws.append(struct.pack("<IIIIIBBH", rank, int(s[a]), int(e[b - 1]), 0, 0, 1, 0, b - a) + body[12 * a:12 * b])
            sec_items.append([rank, int(s[a]), int(e[b - 1]), 0])
        max_raw = max(max_raw, max(len(r) for r in raws))
        sections.extend(_deflate_all(raws) if compress else raws)
    for item, sec in zip(sec_items, sections):
        item[3] = len(sec)
    if not sections:
        raise ValueError("No intervals to write")

    # ---- zoom levels: reductions x4 from ~10x the mean item width, while they still shrink the data ----
    mean_width = max(1, covered // max(1, sum(len(x[1]) for x in per_chrom)))
    zooms = []
    reduction = mean_width * 10
    n_items = sum(len(x[1]) for x in per_chrom)
    prev = n_items
    work = None
    for _ in range(max(0, int(zoom_levels))):
        if reduction > 0xFFFFFFFF // 4:
            break
        if work is None:
            work = np.concatenate([_zoom_records(rank, s, e, v, reduction) for rank, s, e, v in per_chrom])
        else:
            work = _zoom_coarsen(work, reduction)
        if len(work) >= prev or len(work) == 0:
            break
        recs = np.zeros(len(work), dtype=_ZOOM_DT)
        for field in _ZOOM_DT.names:
            recs[field] = work[field]
        zooms.append((reduction, recs))
        prev = len(work)
        if len(work) <= 1:
            break
        reduction *= 4
    for _red, recs in zooms:
        max_raw = max(max_raw, min(len(recs), ZOOM_RECORDS_PER_SLOT) * _ZOOM_DT.itemsize)

    # ---- layout ----
    key_size = max(len(c.encode()) for c, _ in chrom_sizes)
    n_chrom = len(chrom_sizes)
    pos = 64 + 24 * len(zooms)
    total_summary_offset = pos
    pos += 40
    chrom_tree_offset = pos
    chrom_tree = _chrom_tree_bytes(chrom_sizes, ids, key_size, chrom_tree_offset)
    pos += len(chrom_tree)
    full_data_offset = pos
    pos += 8  # section count
    items = np.zeros(len(sections), dtype=_ITEM_DT)
    for i, (rank, s0, e1, size) in enumerate(sec_items):
        items[i] = (rank, s0, rank, e1, pos, size)
        pos += size
    full_index_offset = pos
    index_bytes = _rtree_bytes(items, full_index_offset, ITEMS_PER_SLOT)
    pos += len(index_bytes)
    zoom_blobs = []
    for reduction, recs in zooms:
        data_offset = pos
        blob = bytearray(struct.pack("<I", len(recs)))
        pos += 4
        zitems = []
        slots = [recs[a:a + ZOOM_RECORDS_PER_SLOT] for a in range(0, len(recs), ZOOM_RECORDS_PER_SLOT)]
        comps = [sl.tobytes() for sl in slots]
        if compress:
            comps = _deflate_all(comps)
        for sl, comp in zip(slots, comps):
            # a slot may span chromosomes: its bounding box runs from its first to its last record
            zitems.append((int(sl[0]["chrom"]), int(sl[0]["start"]), int(sl[-1]["chrom"]), int(sl[-1]["end"]), pos, len(comp)))
            blob += comp
            pos += len(comp)
        zi = np.array(zitems, dtype=_ITEM_DT)
        index_offset = pos
        zindex = _rtree_bytes(zi, index_offset, ZOOM_RECORDS_PER_SLOT)
        pos += len(zindex)
        zoom_blobs.append((reduction, data_offset, index_offset, bytes(blob), zindex))

    header = struct.pack("<IHHQQQHHQQIQ", BIGWIG_MAGIC, 4, len(zooms), chrom_tree_offset, full_data_offset, full_index_offset,
                         0, 0, 0, total_summary_offset, max_raw if compress else 0, 0)
    assert len(header) == 64
    out_dir = os.path.dirname(os.path.abspath(path)) or "."
    fd, tmp = tempfile.mkstemp(prefix="consenrich_bigwig_", suffix=".bw", dir=out_dir)
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(header)
            for reduction, data_offset, index_offset, _blob, _zindex in zoom_blobs:
                f.write(struct.pack("<IIQQ", reduction, 0, data_offset, index_offset))
            f.write(struct.pack("<Qdddd", covered, vmin, vmax, vsum, vsumsq))
            f.write(chrom_tree)
            f.write(struct.pack("<Q", len(sections)))
            for sec in sections:
                f.write(sec)
            f.write(index_bytes)
            for _reduction, _do, _io, blob, zindex in zoom_blobs:
                f.write(blob)
                f.write(zindex)
            assert f.tell() == pos, (f.tell(), pos)
            f.write(struct.pack("<I", BIGWIG_MAGIC))  # the end signature UCSC's and libBigWig's writers leave
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def values_as_printed(values) -> np.ndarray:
    """The float32 a bigWig holds for a value that went through the bedGraph text: ``"%.4f" % v`` parsed back
    (io.py:728) and narrowed by pyBigWig.  For a float32 v the product v * 10^4 is exact in double, so its
    half-even ``rint`` is printf's correctly rounded digit string and the quotient by 10^4 the double that string
    parses to -- the track can skip the text without changing a bit of the file."""
    v = np.asarray(values, dtype=np.float32).astype(np.float64)
    return (np.rint(v * 1.0e4) / 1.0e4).astype(np.float32)


def fixed_step_track(chromosome: str, values, *, start0: int = 0, step: int, chrom_size: Optional[int] = None):
    """(chromosome, starts, ends, values) of a fixed-step track for ``write_bigwig``: interval i is
    [start0 + i step, start0 + (i + 1) step), the last one cut at the chromosome end the way the bedGraph rows
    are; values as the bedGraph would have printed them."""
    v = values_as_printed(np.asarray(values).reshape(-1))
    starts = int(start0) + np.arange(len(v), dtype=np.int64) * int(step)
    ends = starts + int(step)
    if chrom_size is not None:
        ends = np.minimum(ends, int(chrom_size))
    return chromosome, starts, ends, v


def convert_bedgraph_to_bigwig(bedgraph_path: str, chrom_sizes, bigwig_path_: str, *, zoom_levels: int = MAX_ZOOM_LEVELS) -> None:
    """``_convertBedGraphToBigWigPyBigWig`` (io.py:633-780) without pyBigWig: the same row validation, in the
    same order and with the same messages, then ``write_bigwig``.  ``chrom_sizes``: a chromosome-sizes file or a
    sequence of (name, size)."""
    sizes_file = chrom_sizes if isinstance(chrom_sizes, (str, os.PathLike)) else "<chrom sizes>"
    sizes = read_chrom_sizes(chrom_sizes) if isinstance(chrom_sizes, (str, os.PathLike)) else [(str(c), int(s)) for c, s in chrom_sizes]
    if len(sizes) == 0:
        raise ValueError(f"No chromosome sizes found in {sizes_file}")
    size_of = dict(sizes)
    rank_of = {c: r for r, (c, _s) in enumerate(sizes)}
    per: Dict[str, Tuple[list, list, list]] = {}
    order: List[str] = []
    seen = False
    last_chrom, last_start, last_end = "", -1, -1
    with open(bedgraph_path, "r", encoding="utf-8") as handle:
        for line_number, line in enumerate(handle, start=1):
            stripped = line.strip()
            if _is_header_line(stripped):
                continue
            parts = stripped.split()
            if len(parts) != 4:
                raise ValueError(f"Malformed bedGraph row {line_number} in {bedgraph_path}: expected 4 columns")
            chrom = str(parts[0])
            if chrom not in size_of:
                raise ValueError(f"Chromosome {chrom} on bedGraph row {line_number} is not present in {sizes_file}")
            try:
                start, end = int(parts[1]), int(parts[2])
            except ValueError as e:
                raise ValueError(f"Invalid bedGraph coordinates on row {line_number} in {bedgraph_path}") from e
            try:
                value = float(parts[3])
            except ValueError as e:
                raise ValueError(f"Invalid bedGraph value on row {line_number} in {bedgraph_path}") from e
            if not np.isfinite(value):
                raise ValueError(f"Non-finite bedGraph value on row {line_number} in {bedgraph_path}")
            if start < 0:
                raise ValueError(f"Negative start coordinate on bedGraph row {line_number}")
            if end <= start:
                raise ValueError(f"End coordinate must be greater than start on bedGraph row {line_number}")
            if end > size_of[chrom]:
                raise ValueError(f"End coordinate {end} on bedGraph row {line_number} exceeds {chrom} size of {size_of[chrom]}")
            if seen:
                if rank_of[chrom] < rank_of[last_chrom] or (chrom == last_chrom and start < last_start):
                    raise ValueError(f"bedGraph input is not sorted at row {line_number}; sort by chromosome sizes order, "
                                     "then start/end")
                if chrom == last_chrom and start < last_end:
                    raise ValueError(f"Overlapping bedGraph interval at row {line_number}")
            if chrom not in per:
                per[chrom] = ([], [], [])
                order.append(chrom)
            s_, e_, v_ = per[chrom]
            s_.append(start)
            e_.append(end)
            v_.append(value)
            seen, last_chrom, last_start, last_end = True, chrom, start, end
    if not seen:
        raise ValueError(f"No bedGraph intervals found in {bedgraph_path}")
    write_bigwig(bigwig_path_, sizes, [(c, np.array(per[c][0]), np.array(per[c][1]), np.array(per[c][2], np.float32)) for c in order],
                 zoom_levels=zoom_levels)


def _is_header_line(stripped: str) -> bool:
    return (not stripped or stripped.startswith("#") or stripped == "track" or stripped.startswith("track ")
            or stripped == "browser" or stripped.startswith("browser "))


def sort_bedgraph_in_place(bedgraph_path: str, chrom_order: Sequence[str]) -> None:
    """``_sortBedGraphInPlace`` with a chromosome order (io.py:879-990): header lines first, rows by chromosome
    rank, start, end (stable), values re-printed as ``%.4f``.  The repair the reference applies to a bedGraph
    that fails its sortedness check before conversion (io.py:562-579)."""
    if not os.path.exists(bedgraph_path) or os.path.getsize(bedgraph_path) == 0:
        return
    headers: List[str] = []
    chroms: List[str] = []
    cols: List[Tuple[int, int, float]] = []
    with open(bedgraph_path, "r", encoding="utf-8") as handle:
        for line_number, line in enumerate(handle, start=1):
            stripped = line.strip()
            if _is_header_line(stripped):
                headers.append(line.rstrip("\n"))
                continue
            parts = stripped.split()
            if len(parts) != 4:
                raise ValueError(f"Malformed bedGraph row {line_number} in {bedgraph_path}: expected 4 columns")
            chroms.append(parts[0])
            cols.append((int(parts[1]), int(parts[2]), float(parts[3])))
    rank_of = {str(c): r for r, c in enumerate(chrom_order)}
    unknown = sorted({c for c in chroms if c not in rank_of})
    if unknown:
        raise ValueError("bedGraph contains chromosomes not present in chromosome order: " + ", ".join(unknown[:5]))
    out_dir = os.path.dirname(os.path.abspath(bedgraph_path)) or "."
    fd, tmp = tempfile.mkstemp(prefix="consenrich_sort_", suffix=".bedGraph", dir=out_dir)
    try:
        with os.fdopen(fd, "w", encoding="utf-8") as out:
            for h in headers:
                out.write(f"{h}\n")
            if cols:
                arr = np.array(cols, dtype=np.float64)
                ranks = np.array([rank_of[c] for c in chroms], dtype=np.int64)
                starts, ends = arr[:, 0].astype(np.int64), arr[:, 1].astype(np.int64)
                order = np.lexsort((ends, starts, ranks))  # stable: ties keep the file's order, as mergesort does
                for i in order.tolist():
                    out.write("%s\t%d\t%d\t%.4f\n" % (chroms[i], starts[i], ends[i], arr[i, 2]))
        os.replace(tmp, bedgraph_path)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def convert_outputs(experiment_name: str, chrom_sizes_file: str, suffixes: Optional[Sequence[str]] = None, *, version: str,
                    delete_bedgraphs: bool = False, directory: str = ".", validated: Sequence[str] = ()) -> List[str]:
    """The loop of ``convertBedGraphToBigWig`` (io.py:530-600) over a run's finished bedGraph files
    (``consenrichOutput_{experiment}_{suffix}.v{version}.bedGraph``): a missing bedGraph is skipped with a
    warning, a missing chromosome-sizes file ends the conversion, a bedGraph not listed in ``validated`` that
    fails the conversion's order / overlap checks is sorted in place first (the reference's fallback,
    io.py:562-579) and converted then; a track that still fails is reported and skipped.  Returns the bigWig
    files written."""
    import warnings
    written: List[str] = []
    sizes = None
    validated_paths = {os.path.abspath(str(p)) for p in validated}
    for suffix in (["state"] if suffixes is None else list(suffixes)):
        bedgraph = os.path.join(directory, f"consenrichOutput_{experiment_name}_{suffix}.v{version}.bedGraph")
        if not os.path.exists(bedgraph):
            warnings.warn(f"bedGraph file {bedgraph} does not exist. Skipping bigWig conversion.")
            continue
        if not os.path.exists(chrom_sizes_file):
            warnings.warn(f"{chrom_sizes_file} does not exist. Skipping bigWig conversion.")
            return written
        if sizes is None:
            sizes = read_chrom_sizes(chrom_sizes_file)
        out = os.path.join(directory, bigwig_path(experiment_name, suffix, version))
        try:
            try:
                convert_bedgraph_to_bigwig(bedgraph, sizes, out)
            except ValueError as first:
                if os.path.abspath(bedgraph) in validated_paths or "not sorted" not in str(first):
                    raise
                warnings.warn(f"bedGraph {bedgraph} failed sorted validation before bigWig conversion; sorting as a "
                              f"fallback:\n{first}")
                sort_bedgraph_in_place(bedgraph, [c for c, _s in sizes])
                convert_bedgraph_to_bigwig(bedgraph, sizes, out)
        except Exception as e:  # the reference logs and moves on to the next track (io.py:589-593)
            warnings.warn(f"bedGraph-->bigWig conversion for {bedgraph} raised:\n{e}\n")
            continue
        if os.path.exists(out) and os.path.getsize(out) > 100:
            written.append(out)
            if delete_bedgraphs:
                os.remove(bedgraph)
    return written


# ------------------------------------------------------------------------------------------
# reader (independent of the writer's bookkeeping: everything is found through the file's own offsets)
# ------------------------------------------------------------------------------------------
def _read_rtree(buf: bytes, offset: int):
    magic, block, count, sc, sb, ec, eb, end_file, per_slot, _ = struct.unpack_from("<IIQIIIIQII", buf, offset)
    if magic != CIRTREE_MAGIC:
        raise ValueError("bad R-tree magic")
    leaves = []

    def walk(pos, box):
        is_leaf, _r, n = struct.unpack_from("<BBH", buf, pos)
        pos += 4
        for _ in range(n):
            if is_leaf:
                a, b, c_, d, off, size = struct.unpack_from("<IIIIQQ", buf, pos)
                pos += 32
                if box is not None and not ((box[0], box[1]) <= (a, b) and (c_, d) <= (box[2], box[3])):
                    raise ValueError("R-tree child outside its parent's bounding box")
                leaves.append((a, b, c_, d, off, size))
            else:
                a, b, c_, d, child = struct.unpack_from("<IIIIQ", buf, pos)
                pos += 24
                if box is not None and not ((box[0], box[1]) <= (a, b) and (c_, d) <= (box[2], box[3])):
                    raise ValueError("R-tree child outside its parent's bounding box")
                walk(child, (a, b, c_, d))
    walk(offset + 48, None)
    if len(leaves) != count:
        raise ValueError(f"R-tree holds {len(leaves)} leaves, header says {count}")
    return {"block": block, "count": count, "bounds": (sc, sb, ec, eb), "end_file": end_file, "items_per_slot": per_slot,
            "leaves": leaves}


def read_bigwig(path: str) -> dict:
    """Parse a bigWig file: chromosomes, total summary, every interval (through the R-index), zoom levels."""
    with open(path, "rb") as handle:
        buf = handle.read()
    (magic, version, n_zoom, chrom_off, data_off, index_off, field_count, defined, autosql, summary_off, uncompress,
     _ext) = struct.unpack_from("<IHHQQQHHQQIQ", buf, 0)
    if magic != BIGWIG_MAGIC:
        raise ValueError("not a bigWig file")
    end_signature = struct.unpack_from("<I", buf, len(buf) - 4)[0] == BIGWIG_MAGIC
    inflate = (lambda b: zlib.decompress(b)) if uncompress else (lambda b: b)
    zoom_headers = [struct.unpack_from("<IIQQ", buf, 64 + 24 * i) for i in range(n_zoom)]
    covered, vmin, vmax, vsum, vsumsq = struct.unpack_from("<Qdddd", buf, summary_off)
    bmagic, block, key_size, val_size, n_chrom, _res = struct.unpack_from("<IIIIQQ", buf, chrom_off)
    if bmagic != BPT_MAGIC or val_size != 8:
        raise ValueError("bad chromosome tree")
    chroms = {}

    def walk_bpt(pos):
        is_leaf, _r, n = struct.unpack_from("<BBH", buf, pos)
        pos += 4
        for _ in range(n):
            key = buf[pos:pos + key_size].rstrip(b"\0").decode()
            pos += key_size
            if is_leaf:
                cid, size = struct.unpack_from("<II", buf, pos)
                chroms[cid] = (key, size)
            else:
                walk_bpt(struct.unpack_from("<Q", buf, pos)[0])
            pos += 8
    walk_bpt(chrom_off + 32)
    n_sections = struct.unpack_from("<Q", buf, data_off)[0]
    index = _read_rtree(buf, index_off)
    intervals = {name: ([], [], []) for name, _ in chroms.values()}
    for a, b, c_, d, off, size in index["leaves"]:
        raw = inflate(buf[off:off + size])
        if uncompress and len(raw) > uncompress:
            raise ValueError("section larger than the header's uncompress buffer")
        cid, s0, e1, step, span, typ, _r, cnt = struct.unpack_from("<IIIIIBBH", raw, 0)
        if typ != 1:
            raise ValueError("only bedGraph-type sections are written here")
        rec = np.frombuffer(raw, dtype=[("s", "<u4"), ("e", "<u4"), ("v", "<f4")], count=cnt, offset=24)
        if (a, b, d) != (cid, int(rec["s"][0]), int(rec["e"][-1])) or c_ != cid or (s0, e1) != (b, d):
            raise ValueError("R-tree leaf does not describe its section")
        name = chroms[cid][0]
        intervals[name][0].append(rec["s"].astype(np.int64))
        intervals[name][1].append(rec["e"].astype(np.int64))
        intervals[name][2].append(rec["v"].copy())
    tracks = {k: tuple(np.concatenate(x) if x else np.zeros(0) for x in v) for k, v in intervals.items()}
    zooms = []
    for reduction, _res, z_data, z_index in zoom_headers:
        n_rec = struct.unpack_from("<I", buf, z_data)[0]
        zi = _read_rtree(buf, z_index)
        recs = [np.frombuffer(inflate(buf[off:off + size]), dtype=_ZOOM_DT) for _a, _b, _c, _d, off, size in zi["leaves"]]
        recs = np.concatenate(recs) if recs else np.zeros(0, dtype=_ZOOM_DT)
        if len(recs) != n_rec:
            raise ValueError("zoom record count mismatch")
        zooms.append({"reduction": reduction, "records": recs})
    return {"version": version, "chroms": [chroms[i] for i in sorted(chroms)], "sections": int(n_sections),
            "summary": {"bases_covered": covered, "min": vmin, "max": vmax, "sum": vsum, "sum_squares": vsumsq},
            "tracks": tracks, "index": index, "zooms": zooms, "uncompress_buf_size": uncompress,
            "end_signature": end_signature, "field_count": field_count, "defined_field_count": defined, "autosql_offset": autosql}


def query_bigwig(path: str, chromosome: str, start: int, end: int):
    """Intervals of ``chromosome`` overlapping [start, end), found the way a genome browser finds them: chromosome
    id through the B+ tree, then a descent of the R-index that follows only the children whose bounding box
    overlaps the query, then the matching sections.  Returns (starts, ends, values)."""
    with open(path, "rb") as handle:
        buf = handle.read()
    (magic, _version, _n_zoom, chrom_off, _data_off, index_off, _fc, _dfc, _as, _summ, uncompress,
     _ext) = struct.unpack_from("<IHHQQQHHQQIQ", buf, 0)
    if magic != BIGWIG_MAGIC:
        raise ValueError("not a bigWig file")
    _bm, _block, key_size, _vs, _count, _res = struct.unpack_from("<IIIIQQ", buf, chrom_off)
    key = chromosome.encode().ljust(key_size, b"\0")
    pos, cid = chrom_off + 32, None
    while cid is None:
        is_leaf, _r, n = struct.unpack_from("<BBH", buf, pos)
        pos += 4
        nxt = None
        for i in range(n):
            k = buf[pos:pos + key_size]
            if is_leaf:
                if k == key:
                    cid = struct.unpack_from("<I", buf, pos + key_size)[0]
                    break
            elif k <= key:
                nxt = struct.unpack_from("<Q", buf, pos + key_size)[0]
            pos += key_size + 8
        if cid is None:
            if is_leaf or nxt is None:
                raise KeyError(chromosome)
            pos = nxt
    lo, hi = (cid, int(start)), (cid, int(end))
    blocks = []

    def descend(pos):
        is_leaf, _r, n = struct.unpack_from("<BBH", buf, pos)
        pos += 4
        for _ in range(n):
            a, b, c_, d = struct.unpack_from("<IIII", buf, pos)
            overlaps = (a, b) < hi and (c_, d) > lo
            if is_leaf:
                if overlaps:
                    blocks.append(struct.unpack_from("<QQ", buf, pos + 16))
                pos += 32
            else:
                if overlaps:
                    descend(struct.unpack_from("<Q", buf, pos + 16)[0])
                pos += 24
    descend(index_off + 48)
    out_s, out_e, out_v = [], [], []
    for off, size in blocks:
        raw = zlib.decompress(buf[off:off + size]) if uncompress else buf[off:off + size]
        sec_chrom, _s0, _e1, _step, _span, _typ, _r, cnt = struct.unpack_from("<IIIIIBBH", raw, 0)
        if sec_chrom != cid:
            continue
        rec = np.frombuffer(raw, dtype=[("s", "<u4"), ("e", "<u4"), ("v", "<f4")], count=cnt, offset=24)
        keep = (rec["s"] < end) & (rec["e"] > start)
        out_s.append(rec["s"][keep].astype(np.int64))
        out_e.append(rec["e"][keep].astype(np.int64))
        out_v.append(rec["v"][keep])
    if not out_s:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    return np.concatenate(out_s), np.concatenate(out_e), np.concatenate(out_v)
