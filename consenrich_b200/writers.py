"""bedGraph output of the path: the rows ``consenrich.consenrich.main`` appends per
chromosome and track with pandas (consenrich.py:9797-9805)::

    df[["Chromosome", "Start", "End", col]].to_csv(path, sep="\\t", header=False, index=False,
                                                   mode="w" if first else "a", float_format="%.4f",
                                                   lineterminator="\\n")

formatted on the device (csrc/writer_kernels.cu), byte-identical to that call for float32 tracks -- correct
``%.4f`` rounding of the exact value, ``-0.0000``, empty field for NaN, ``inf`` / ``-inf`` included.
File names follow the reference (consenrich.py:9790).

bigWig conversion stays what it is in the reference: a call into the third-party pyBigWig on the finished
bedGraph (io.py:633-780); it is not part of this library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .native import _ctx

__all__ = ["bedgraph_chunk", "write_bedgraph_chunk", "bedgraph_path"]


def bedgraph_path(experiment_name: str, suffix: str, version: str) -> str:
    """consenrichOutput_{experimentName}_{suffix}.v{version}.bedGraph (consenrich.py:9790)."""
    return f"consenrichOutput_{experiment_name}_{suffix}.v{version}.bedGraph"


def bedgraph_chunk(chromosome: str, values, starts=None, ends=None, *, start0: int = 0, step: int = 0,
                   end_clip: int = 0) -> bytes:
    """The text of one chromosome's rows.  ``values``: float32 [n], or [n, d] (column 0 -- the level of a
    state array -- is printed).  Intervals: ``starts`` / ``ends`` (int64 [n]), or uniform ``start0 + k step``
    with ``end = start + step`` (clipped to ``end_clip`` when given)."""
    v = np.asarray(values)
    if v.dtype != np.float32:
        raise TypeError("values must be float32 (the reference writes float32 tracks; a float64 column prints differently)")
    if v.ndim == 2:
        stride = v.shape[1]
        v = np.ascontiguousarray(v)
    elif v.ndim == 1:
        stride = 1
        v = np.ascontiguousarray(v)
    else:
        raise ValueError("values must have shape (n,) or (n, d)")
    n = v.shape[0]
    if (starts is None) != (ends is None):
        raise ValueError("starts and ends are given together or not at all")
    s = e = None
    if starts is not None:
        s = np.ascontiguousarray(starts, dtype=np.int64).reshape(-1)
        e = np.ascontiguousarray(ends, dtype=np.int64).reshape(-1)
        if s.shape[0] != n or e.shape[0] != n:
            raise ValueError("starts / ends length must match values")
    elif step <= 0 and n > 0:
        raise ValueError("uniform intervals need a positive step")
    name = chromosome.encode()
    if not 0 < len(name) <= 32:
        raise ValueError("chromosome name must be 1..32 bytes")
    if n == 0:
        return b""
    ctx = _ctx()
    text, nbytes = C.c_void_p(), C.c_int64(0)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(ctx._lib.cb200_host_bedgraph_chunk(ctx.handle, name, n, ptr(s), ptr(e), int(start0), int(step), int(end_clip),
                                                  ptr(v), int(stride), C.byref(text), C.byref(nbytes)))
    return C.string_at(text.value, nbytes.value)  # copied out: the context's buffer is reused by the next call


def write_bedgraph_chunk(path: str, chromosome: str, values, starts=None, ends=None, *, first: bool, start0: int = 0,
                         step: int = 0, end_clip: int = 0) -> int:
    """Append (or, for the first chromosome, create) the chunk, like the reference's ``mode="w" if c_ == 0
    else "a"``; returns the bytes written."""
    text = bedgraph_chunk(chromosome, values, starts, ends, start0=start0, step=step, end_clip=end_clip)
    with open(path, "wb" if first else "ab") as f:
        f.write(text)
    return len(text)
