import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import tiler_b200
        return tiler_b200.api.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def tm():
    import tiler_b200
    if not _have_gpu():
        pytest.fail("no sm_100 device / libtm_gpu.so not usable: GPU tests must run on a B200 (no CPU fallback)")
    return tiler_b200.api


def rand_tiles(n, seed, smooth=True):
    """Random RGB tiles [n,64] int32: smooth gradients + noise (realistic feature spread) or pure noise."""
    rng = np.random.default_rng(seed)
    if not smooth:
        c = rng.integers(0, 256, size=(n, 64, 3))
    else:
        base = rng.integers(0, 256, size=(n, 1, 3)).astype(np.float64)
        gx = rng.normal(0, 12, size=(n, 1, 3)); gy = rng.normal(0, 12, size=(n, 1, 3))
        x = (np.arange(64) % 8)[None, :, None]; y = (np.arange(64) // 8)[None, :, None]
        c = np.clip(base + gx * x + gy * y + rng.normal(0, 6, size=(n, 64, 3)), 0, 255).astype(np.int64)
    return (c[..., 0] | (c[..., 1] << 8) | (c[..., 2] << 16)).astype(np.int32)


def rand_palettes(n_pal, pal_size, seed, n_null=0):
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 256, size=(n_pal, pal_size, 3))
    p = (c[..., 0] | (c[..., 1] << 8) | (c[..., 2] << 16)).astype(np.int32)
    if n_null:
        p[:, pal_size - n_null:] = np.int32(-65281)
    return p
