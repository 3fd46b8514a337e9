"""A chromosome split into contiguous bin ranges: the shard entry points of the C ABI
(aggregate -> gathered prefix -> scan with a carried state, and the mirror image for the smoother)
must reproduce the unsharded sweep.  All shards run on ONE GPU, one after the other
(consenrich_b200.sharding.run_split_local); the multi-process plumbing is covered on the CPU by
tests/test_sharding_host.py."""
import numpy as np
import pytest

from conftest import synth_tracks
from parity_util import assert_sweep_tracks_close
from test_gpu_parity import Q0, F, _sweep, _weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [2, 1])
@pytest.mark.parametrize("shards,n", [(2, 40_000), (3, 100_003), (5, 7_777)])
def test_split_chromosome_matches_unsharded_oracle(oracle, dim, shards, n):
    import torch

    from consenrich_b200 import sharding
    from consenrich_b200.device import TrackSweep, make_model

    m = 6
    data, munc = synth_tracks(77 + n, m, n, masked_frac=0.02)
    lam, kap, qs = _weights(np.random.default_rng(n), n)
    want = _sweep(oracle, dim, data, munc, lam, kap, qs)

    dev = torch.device("cuda", 0)
    model = make_model(dim, F, Q0, 0.25, 1000.0, 1e-4, lam_bounds=(0.25, 4.0), kap_bounds=(5e-3, 5e3),
                       return_nll=True, use_lambda=True, use_kappa=True, use_qscale=True)
    ranges = sharding.split_ranges(n, shards, align=512)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a < b for a, b in ranges)
    backends = []
    for a, b in ranges:
        nb = b - a
        ld = (nb + 31) // 32 * 32
        d_dev = torch.zeros((m, ld), dtype=torch.float32, device=dev)
        v_dev = torch.ones((m, ld), dtype=torch.float32, device=dev)
        d_dev[:, :nb] = torch.from_numpy(data[:, a:b]).to(dev)
        v_dev[:, :nb] = torch.from_numpy(munc[:, a:b]).to(dev)
        vec = lambda x: torch.from_numpy(np.ascontiguousarray(x[a:b])).to(dev)
        ts = TrackSweep(m, nb, dim, 0, residuals=True)
        backends.append(sharding.DeviceShard(ts, model, d_dev, v_dev, ld, vec(lam), vec(kap), vec(qs)))
    sums = sharding.run_split_local(backends).cpu().numpy()
    torch.cuda.synchronize()

    cat = lambda name: np.concatenate([getattr(b.ts, name).cpu().numpy() for b in backends])
    got = dict(xf=cat("xf"), Pf=cat("Pf"), D=cat("D"), xs=cat("xs"), Ps=cat("Ps"), res=cat("resid"))
    Qf = cat("Qf")
    lag = cat("lag")  # every shard holds n_local rows; row n-1 of the last one does not exist
    close = assert_sweep_tracks_close
    close(got["xf"], want["xf"], "stateForward")
    close(got["Pf"], want["Pf"], "stateCovarForward", scale="component")
    np.testing.assert_array_equal(Qf[: n - 1], want["Qf"][: n - 1])
    close(got["D"], want["D"], "vectorD")
    close(got["xs"], want["xs"], "stateSmoothed")
    close(got["Ps"], want["Ps"], "stateCovarSmoothed", scale="component")
    close(lag[: n - 1], want["lag"], "lagCovSmoothed", scale="component")
    close(got["res"], want["res"], "postFitResiduals")
    assert abs(sums[1] - want["nll"]) <= 2e-6 * abs(want["nll"])
    assert abs(np.float32(sums[0] / n) - want["phi"]) <= 1e-4 * max(abs(want["phi"]), 1e-3)


@pytest.mark.parametrize("shards,n,qscale", [(2, 40_000, False), (3, 100_003, True), (5, 70_001, False), (8, 300_017, True)])
def test_split_chromosome_ecm_matches_unsharded_oracle(oracle, shards, n, qscale):
    """cfixedBackgroundECM on a chromosome split into contiguous ranges (C ABI cb200_split_*; one payload per pass
    is all the shards exchange): every shard on ONE GPU here, one context each, the collectives replaced by a
    stack (sharding.LocalGather); tests/test_sharding_host.py runs the same loop over a 2-process gloo group and
    tools/split_ecm_nccl.py over NCCL.  Must reproduce the unsharded reference: kappa across the boundaries
    included (cconsenrich.pyx:8244-8298 couples intervals k and k+1 of neighbouring shards)."""
    import torch

    from consenrich_b200 import _lib, sharding
    from consenrich_b200.device import make_model
    from test_gpu_parity import ECM_TOL, _ecm

    m = 5
    data, munc = synth_tracks(900 + n, m, n, masked_frac=0.02)
    rng = np.random.default_rng(n)
    qs = (0.5 + rng.random(n)).astype(np.float32)
    qs[0] = 1.0
    warm = np.exp(rng.normal(0, 0.7, n)).astype(np.float32)
    opts = dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2, ECM_robustTNu=6.0,
                ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3,
                processPrecExpInit=warm, processQScale=qs if qscale else None)
    want = _ecm(oracle, 2, data, munc, **opts)

    dev = torch.device("cuda", 0)
    stream = int(torch.cuda.current_stream(dev).cuda_stream) or 1
    model = make_model(2, F, Q0, 0.0, 1000.0, 1e-4, kap_bounds=(5e-3, 5e3), return_nll=True, use_kappa=True,
                       use_qscale=qscale)
    ranges = sharding.split_ranges(n, shards, align=512)
    parts = []
    for r, (a, b) in enumerate(ranges):
        nb = b - a
        ld = (nb + 31) // 32 * 32
        d_dev = torch.zeros((m, ld), dtype=torch.float32, device=dev)
        v_dev = torch.ones((m, ld), dtype=torch.float32, device=dev)
        d_dev[:, :nb] = torch.from_numpy(data[:, a:b]).to(dev)
        v_dev[:, :nb] = torch.from_numpy(munc[:, a:b]).to(dev)
        kap = torch.from_numpy(np.clip(warm[a:b], 5e-3, 5e3)).to(dev)
        qd = torch.from_numpy(np.ascontiguousarray(qs[a:b])).to(dev) if qscale else None
        parts.append(sharding.EcmShard(_lib.Context(0, stream), model, 6.0, d_dev, v_dev, ld, nb, kap, qd, r, shards))
    diag = sharding.split_ecm(parts, sharding.LocalGather(), max_iters=3, inner_iters=2, rtol=0.0)
    torch.cuda.synchronize()
    cat = lambda name: np.concatenate([getattr(p, name).cpu().numpy() for p in parts])
    assert diag["iters_done"] == want[0]
    assert abs(diag["final_nll"] - want[1]) <= ECM_TOL["nll"] * abs(want[1])
    from parity_util import assert_tracks_close
    assert_tracks_close(cat("xs"), want[2], "split stateSmoothed", **ECM_TOL["state"])
    assert_tracks_close(cat("Ps"), want[3], "split stateCovarSmoothed", scale="component", **ECM_TOL["cov"])
    assert_tracks_close(cat("lag")[: n - 1], want[4], "split lagCovSmoothed", scale="component", **ECM_TOL["cov"])
    assert_tracks_close(cat("resid"), want[5], "split residuals", **ECM_TOL["state"])
    assert_tracks_close(cat("kap"), want[7], "split kappa", **ECM_TOL["mult"])
    assert cat("kap")[0] == 1.0


def test_split_chromosome_ecm_replayed_from_cuda_graphs(oracle):
    """The passes of split_ecm captured once into CUDA graphs (stream capture on the shards' stream) and replayed:
    the first call captures while it runs, the second only replays -- both must give the unsharded reference."""
    import torch

    from consenrich_b200 import _lib, sharding
    from consenrich_b200.device import make_model
    from parity_util import assert_tracks_close
    from test_gpu_parity import ECM_TOL, _ecm

    m, n, shards = 4, 120_011, 3
    data, munc = synth_tracks(4321, m, n, masked_frac=0.02)
    opts = dict(ECM_fixedBackgroundIters=3, ECM_fixedBackgroundRtol=0.0, t_innerIters=2, ECM_robustTNu=8.0,
                ECM_useObsPrecisionReweighting=False, procPrecisionMultiplierMin=5e-3, procPrecisionMultiplierMax=5e3)
    want = _ecm(oracle, 2, data, munc, **opts)
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(dev)  # capture needs a non-default stream, and the library must launch on the same one
    with torch.cuda.stream(side):
        model = make_model(2, F, Q0, 0.0, 1000.0, 1e-4, kap_bounds=(5e-3, 5e3), return_nll=True, use_kappa=True)
        parts = []
        for r, (a, b) in enumerate(sharding.split_ranges(n, shards, align=512)):
            nb = b - a
            ld = (nb + 31) // 32 * 32
            d_dev = torch.zeros((m, ld), dtype=torch.float32, device=dev)
            v_dev = torch.ones((m, ld), dtype=torch.float32, device=dev)
            d_dev[:, :nb] = torch.from_numpy(data[:, a:b]).to(dev)
            v_dev[:, :nb] = torch.from_numpy(munc[:, a:b]).to(dev)
            kap = torch.ones(nb, dtype=torch.float32, device=dev)
            parts.append(sharding.EcmShard(_lib.Context(0, int(side.cuda_stream)), model, 8.0, d_dev, v_dev, ld, nb, kap,
                                           None, r, shards))
        graphs = {}
        for attempt in ("captured", "replayed"):
            for p in parts:
                p.kap.fill_(1.0)
            diag = sharding.split_ecm(parts, sharding.LocalGather(), max_iters=3, inner_iters=2, rtol=0.0, graphs=graphs)
            side.synchronize()
            cat = lambda name: np.concatenate([getattr(p, name).cpu().numpy() for p in parts])
            assert diag["iters_done"] == want[0] and abs(diag["final_nll"] - want[1]) <= ECM_TOL["nll"] * abs(want[1]), attempt
            assert_tracks_close(cat("xs"), want[2], f"{attempt} stateSmoothed", **ECM_TOL["state"])
            assert_tracks_close(cat("Ps"), want[3], f"{attempt} stateCovarSmoothed", scale="component", **ECM_TOL["cov"])
            assert_tracks_close(cat("kap"), want[7], f"{attempt} kappa", **ECM_TOL["mult"])
        assert len(graphs) >= 4  # forward (plain, NLL + store, NLL only), backward (kappa, publish)
