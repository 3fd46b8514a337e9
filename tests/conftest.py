"""pytest configuration: markers, import path, shared synthetic-input helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def synth_tracks(seed, m, n, masked_frac=0.0):
    """Seeded synthetic [m x n] count/variance matrices (SURVEY 8d generator, scaled down)."""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    x = 0.3 * np.sin(2 * np.pi * k / max(n, 8) * 3.0)
    for _ in range(max(1, n // 400)):
        c, w, h = rng.integers(0, n), rng.uniform(3, 40), rng.uniform(0.5, 4.0)
        x = x + h * np.exp(-0.5 * ((k - c) / w) ** 2)
    v0 = rng.uniform(0.05, 0.3, size=(m, 1))
    munc = (v0 * (1.0 + np.abs(x))[None, :] * rng.uniform(0.5, 1.5, size=(m, n))).astype(np.float32)
    data = (x[None, :] + rng.normal(0, 0.05, size=(m, 1))
            + rng.normal(size=(m, n)) * np.sqrt(munc)).astype(np.float32)
    if masked_frac > 0:
        munc[rng.random((m, n)) < masked_frac] = np.float32(1.0e30)
    return np.ascontiguousarray(data), np.ascontiguousarray(munc)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
