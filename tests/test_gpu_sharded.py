"""GPU: measurement-sharded snapshot gradient (config 5).  Single process: the sum of the per-shard
partial gradients equals the full gradient; with 2 visible GPUs the NCCL path is exercised too."""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


def test_partial_gradients_sum_to_full(cuda):
    from pnp_svrg_b200.problems import CSMRI
    img = synth_image(128, 128, 1)
    z = np.random.default_rng(0).uniform(0, 1, 128 * 128)
    np.random.seed(0)
    full = CSMRI(image=img, H=128, W=128, sample_prob=0.3, snr=20.)
    g = full.grad_full(z)
    for world in (2, 4):
        acc = np.zeros_like(g)
        cnt = 0
        for r in range(world):
            np.random.seed(0)
            p = CSMRI(image=img, H=128, W=128, sample_prob=0.3, snr=20., shard=(r, world))
            assert np.array_equal(p.mask, full.mask) and p.M0 == full.M0        # replicas of the same problem
            acc += p.grad_full(z)                                                # partial (divided by the GLOBAL M0)
            cnt += p._shard_count
        assert cnt == full.M0
        assert rel_l2(acc, g) < 2e-6
