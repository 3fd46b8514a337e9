"""Small calls of every drop-in function at three sizes (a quick whole-surface run on a GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import consenrich_b200 as cb  # noqa: E402
from conftest import synth_tracks  # noqa: E402

F = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
Q0 = np.array([[2e-3, 0.0], [0.0, 4e-4]], np.float32)
rng = np.random.default_rng(0)
for m, n in ((3, 7), (5, 1200), (4, 70_001)):
    data, munc = synth_tracks(n, m, n, masked_frac=0.02)
    bm = np.zeros(n, np.int32)
    for dim in (2, 1):
        kw = dict(matrixData=data, matrixPluginMuncInit=munc, matrixQ0=Q0, intervalToBlockMap=bm, blockCount=1,
                  stateInit=0.0, stateCovarInit=1000.0, returnIntermediates=True, ECM_fixedBackgroundIters=2,
                  ECM_fixedBackgroundRtol=0.0)
        out = cb.cfixedBackgroundECM(matrixF=F, **kw) if dim == 2 else cb.cfixedBackgroundECMLevel(**kw)
        assert np.isfinite(out[1])
    cb.sweep(data, munc, F, Q0, 0.0, 1000.0)
    w, rhs, sup = cb.cbackgroundWeightedStatsWithSupport(data, 1.0 / np.maximum(munc, 1e-3))
    x = cb.csolveZeroCenteredBackground(w + 1.0, rhs, 64.0, True, lamFirst=1.0)
    assert abs(x.sum()) < 1e-6 * (np.abs(x).sum() + 1)
    le = (np.abs(data) + 0.01).astype(np.float32)
    for win in (1, 9, 300):
        cb.cMuncSmoothDenseLocalEvidence(le, win, excludeMask=(rng.random(n) < 0.1).astype(np.uint8))
    cb.cFinalizeMuncEBTrack(le[0], le[1 % m], np.where(rng.random(n) < 0.3, np.nan, 0.1).astype(np.float32),
                            nuLocal=3.0, nuPrior=2.0, varianceFloor=1e-3, varianceCap=4.0)
    cb.cMuncObservationMomentSeedPass(data, np.minimum(munc, 10.0), data[0] * 0.5, np.full(n, 0.1, np.float32),
                                      countFloor=np.full((m, n), 0.05, np.float32),
                                      activeMask=(rng.random((m, n)) < 0.9).astype(np.uint8))
print("probe ok")
