import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def synth_image(H, W, seed=0):
    """Deterministic uint8 test image (smooth blobs + edges + texture); the reference's Set12
    PNGs do not travel to the GPU box."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W))
    for _ in range(6):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        s = rng.uniform(0.05, 0.3) * min(H, W)
        img += rng.uniform(0.3, 1.0) * np.exp(-((y - cy) ** 2 + (x - cx) ** 2) / (2 * s * s))
    img += 0.4 * (x / W > rng.uniform(0.3, 0.7)) + 0.3 * (y / H > rng.uniform(0.3, 0.7))
    img += 0.1 * np.sin(2 * np.pi * x / rng.uniform(4, 16)) * np.cos(2 * np.pi * y / rng.uniform(4, 16))
    img += 0.03 * rng.standard_normal((H, W))
    img = (img - img.min()) / (img.max() - img.min())
    return np.round(img * 255).astype(np.uint8)


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    a = a.astype(np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = b.astype(a.dtype)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda')
