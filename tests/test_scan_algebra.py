"""CPU check of the scan algebra the sm_100a kernels use (consenrich_b200/csrc/ssm_math.cuh).

tests/emul/scan_emul.cpp replays the kernels' tiling (thread chunks, Kogge-Stone tile scan,
look-back over tile aggregates with randomised windows, reference-ordered replay) on the CPU
from the same header; here it is compared with the oracle.  Tolerances are the ones the GPU
parity tests use (tests/test_gpu_parity.py): the scan re-associates float64 arithmetic and
the reference rounds its carried state to float32 every bin, so agreement is to float32
resolution of each track's own scale, not bit-exact.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, synth_tracks
from parity_util import assert_tracks_close

F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
_EMUL_SRC = os.path.join(ROOT, "tests", "emul", "scan_emul.cpp")
_EMUL_SO = os.path.join(ROOT, "tests", "emul", "_build", "libscan_emul.so")


class _P(C.Structure):
    _fields_ = [("F", C.c_double * 4), ("Q0", C.c_double * 4), ("state_init", C.c_double),
                ("cov_init", C.c_double), ("lam_min", C.c_double), ("lam_max", C.c_double),
                ("kap_min", C.c_double), ("kap_max", C.c_double), ("use_lambda", C.c_int32),
                ("use_kappa", C.c_int32), ("use_qscale", C.c_int32), ("return_nll", C.c_int32),
                ("store_nll_in_d", C.c_int32), ("do_store", C.c_int32), ("chunk", C.c_int32),
                ("tile_chunks", C.c_int32), ("canon", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64)]


@pytest.fixture(scope="module")
def emul():
    hdr = os.path.join(ROOT, "consenrich_b200", "csrc", "ssm_math.cuh")
    if (not os.path.exists(_EMUL_SO)) or os.path.getmtime(_EMUL_SO) < max(os.path.getmtime(_EMUL_SRC),
                                                                           os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(_EMUL_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off",
                               _EMUL_SRC, "-o", _EMUL_SO])
    return C.CDLL(_EMUL_SO)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fold(emul, data, munc, pad):
    m, n = data.shape
    S = [np.empty(n, np.float64) for _ in range(4)]
    emul.emul_fold(_ptr(data), _ptr(munc), C.c_int64(m), C.c_int64(n), C.c_int64(n),
                   C.c_double(float(np.float32(pad))), *map(_ptr, S))
    return S


def _params(Q0, state_init, chunk, tile_chunks, seed, lam, kap, qs, bounds, return_nll, nll_in_d, canon=0):
    p = _P()
    p.F[:] = [1.0, 1.0, 0.0, 1.0]
    p.Q0[:] = [float(Q0[0, 0]), float(Q0[0, 1]), float(Q0[1, 0]), float(Q0[1, 1])]
    p.state_init, p.cov_init = float(np.float32(state_init)), 1000.0
    p.lam_min, p.lam_max, p.kap_min, p.kap_max = [float(np.float32(b)) for b in bounds]
    p.use_lambda, p.use_kappa, p.use_qscale = int(lam is not None), int(kap is not None), int(qs is not None)
    p.return_nll, p.store_nll_in_d, p.do_store = int(return_nll), int(nll_in_d), 1
    p.chunk, p.tile_chunks, p.seed = chunk, tile_chunks, seed
    p.canon = int(canon)
    return p


CASES = [
    # m, n, masked, weights, chunk (bins per thread run), tile_chunks (runs per tile), canonical-F code path
    (3, 257, 0.0, False, 4, 8, 0),
    (10, 5000, 0.05, True, 8, 32, 1),
    (5, 1, 0.0, False, 8, 32, 1),
    (5, 2, 0.0, True, 8, 32, 0),
    (25, 3001, 0.3, True, 16, 4, 1),
    (2, 20000, 0.0, True, 8, 128, 0),
    (10, 40000, 0.02, True, 32, 128, 1),   # long runs: 4 sub-steps of 8 per thread
    (4, 70000, 0.0, True, 64, 128, 1),     # 8 sub-steps
]


@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("case", CASES)
def test_emulated_scan_matches_oracle(oracle, emul, dim, case):
    m, n, masked, weights, chunk, tile_chunks, canon = case
    data, munc = synth_tracks(1000 + n, m, n, masked_frac=masked)
    rng = np.random.default_rng(n)
    Q0 = np.array([[2e-3, 0.0], [0.0, 1e-4]], np.float32)
    lam = kap = qs = None
    bounds = (0.25, 4.0, 5e-3, 5e3)
    if weights:
        lam = (0.1 + 5 * rng.random(n)).astype(np.float32)
        kap = np.exp(rng.normal(0, 2, n)).astype(np.float32)
        qs = (0.5 + rng.random(n)).astype(np.float32)
        qs[0] = 1.0
    bm = np.zeros(n, np.int32)
    kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=bm, blockCount=1,
              stateInit=0.25, stateCovarInit=1000.0, pad=1e-4, returnNLL=True, storeNLLInD=False,
              lambdaExp=lam, processPrecExp=kap, processQScale=qs, obsPrecisionMultiplierMin=bounds[0],
              obsPrecisionMultiplierMax=bounds[1], procPrecisionMultiplierMin=bounds[2],
              procPrecisionMultiplierMax=bounds[3])
    want = dict(xf=np.empty((n, dim), np.float32), Pf=np.empty((n, dim, dim), np.float32),
                Qf=np.zeros((n, dim, dim), np.float32), D=np.empty(n, np.float32))
    st = dict(stateForward=want["xf"], stateCovarForward=want["Pf"], pNoiseForward=want["Qf"], vectorD=want["D"])
    if dim == 2:
        r = oracle.cforwardPass(matrixF=F, **kw, **st)
        b = oracle.cbackwardPass(matrixData=data, matrixF=F, stateForward=want["xf"], stateCovarForward=want["Pf"],
                                 pNoiseForward=want["Qf"])
    else:
        r = oracle.cforwardPassLevel(**kw, **st)
        b = oracle.cbackwardPassLevel(matrixData=data, stateForward=want["xf"], stateCovarForward=want["Pf"],
                                      pNoiseForward=want["Qf"])
    S = _fold(emul, data, munc, 1e-4)
    p = _params(Q0, 0.25, chunk, tile_chunks, 12345 + n, lam, kap, qs, bounds, True, False, canon)
    got = dict(xf=np.empty((n, dim), np.float32), Pf=np.empty((n, dim, dim), np.float32),
               Qf=np.zeros((n, dim, dim), np.float32), D=np.empty(n, np.float32))
    sd, snll = C.c_double(), C.c_double()
    fwd = emul.emul_forward2 if dim == 2 else emul.emul_forward1
    fwd(*map(_ptr, S), C.c_int64(m), C.c_int64(n), _ptr(lam), _ptr(kap), _ptr(qs), C.byref(p),
        _ptr(got["D"]), _ptr(got["xf"]), _ptr(got["Pf"]), _ptr(got["Qf"]), C.byref(sd), C.byref(snll))
    assert_tracks_close(got["xf"], want["xf"], "stateForward")
    assert_tracks_close(got["Pf"], want["Pf"], "stateCovarForward", scale="component")
    np.testing.assert_array_equal(got["Qf"][: n - 1], want["Qf"][: n - 1])
    assert_tracks_close(got["D"], want["D"], "vectorD")
    assert abs(snll.value - r[3]) <= 1e-7 * max(abs(r[3]), 1.0)
    if dim == 1:  # the level filter carries float64 in the reference too: float32 outputs identical
        np.testing.assert_array_equal(got["xf"], want["xf"])
        np.testing.assert_array_equal(got["Pf"], want["Pf"])
    assert abs(np.float32(sd.value / n) - np.float32(r[0])) <= 1e-5 * max(abs(r[0]), 1e-3)

    # smoother: feed BOTH the oracle's forward outputs (isolates the reverse scan)
    xs, Ps = np.empty((n, dim), np.float32), np.empty((n, dim, dim), np.float32)
    lag = np.zeros((max(n - 1, 1), dim, dim), np.float32)
    if dim == 2:
        Fd = (C.c_double * 4)(1.0, 1.0, 0.0, 1.0)
        emul.emul_backward2(C.c_int64(n), Fd, _ptr(want["xf"]), _ptr(want["Pf"]), _ptr(want["Qf"]), C.byref(p),
                            _ptr(xs), _ptr(Ps), _ptr(lag), C.c_int64(lag.shape[0]))
    else:
        emul.emul_backward1(C.c_int64(n), _ptr(want["xf"]), _ptr(want["Pf"]), _ptr(want["Qf"]), C.byref(p),
                            _ptr(xs), _ptr(Ps), _ptr(lag), C.c_int64(lag.shape[0]))
    assert_tracks_close(xs, b[0], "stateSmoothed")
    assert_tracks_close(Ps, b[1], "stateCovarSmoothed", scale="component")
    if n > 1:
        assert_tracks_close(lag, b[2], "lagCovSmoothed", scale="component")


def test_canonical_smoothing_element_shortcut(emul):
    """smo2_from_filtered_canon (what the lean forward replay composes per bin) is the same element as
    rts2_gain + smo2_from_rts for F = [[1, f], [0, 1]] and symmetric P, Q.  The element is ill-conditioned
    when P^- is (J = P F^T (P^-)^-1, L = P - J P^- J^T cancels), so the two routes are both measured against
    an extended-precision evaluation: the shortcut must not be less accurate than the route it replaces."""
    rng = np.random.default_rng(12)
    ld = np.longdouble
    worse = []
    for _ in range(2000):
        dF = float(rng.choice([1.0, 0.5, 2.0]))
        a = rng.normal(size=(2, 2))
        P = a @ a.T * float(10.0 ** rng.uniform(-6, 3)) + np.eye(2) * 1e-9
        q = rng.normal(size=(2, 2))
        Q = q @ q.T * float(10.0 ** rng.uniform(-8, 0)) + np.eye(2) * 1e-12
        xP = np.array([rng.normal(), rng.normal(), P[0, 0], P[0, 1], P[1, 1]])
        xP[2:] = xP[2:].astype(np.float32)  # the replay feeds float32 values
        Qv = np.array([Q[0, 0], Q[0, 1], Q[1, 1]]).astype(np.float32).astype(np.float64)
        out = np.empty(18)
        emul.emul_smoothing_element(C.c_double(dF), _ptr(xP), _ptr(Qv), _ptr(out))
        Fm = np.array([[1, dF], [0, 1]], ld)
        Pm = np.array([[xP[2], xP[3]], [xP[3], xP[4]]], ld)
        Qm = np.array([[Qv[0], Qv[1]], [Qv[1], Qv[2]]], ld)
        x = np.array([xP[0], xP[1]], ld)
        PP = Fm @ Pm @ Fm.T + Qm
        det = PP[0, 0] * PP[1, 1] - PP[0, 1] * PP[1, 0]
        inv = np.array([[PP[1, 1], -PP[0, 1]], [-PP[1, 0], PP[0, 0]]], ld) / det
        J = Pm @ Fm.T @ inv
        g = x - J @ (Fm @ x)
        Lm = Pm - J @ PP @ J.T
        want = np.array([J[0, 0], J[0, 1], J[1, 0], J[1, 1], g[0], g[1], Lm[0, 0], Lm[0, 1], Lm[1, 1]], ld)
        sc = np.array([1.0] * 4 + [1.0 + abs(xP[0]) + abs(xP[1])] * 2 + [xP[2], max(abs(xP[3]), xP[2] * 1e-3), xP[4]])
        e_general = float(np.max(np.abs(out[:9] - want) / sc))
        e_canon = float(np.max(np.abs(out[9:] - want) / sc))
        worse.append(e_canon - 4.0 * e_general)
        assert e_canon <= 4.0 * e_general + 1e-13, (e_canon, e_general)
