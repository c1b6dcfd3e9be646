"""GPU: the batched sweep engine equals the per-problem engine."""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


def test_batched_equals_single_full_batch(cuda):
    """sample_prob = 1 and B = M0 make the minibatch the whole support whatever the sampler draws, so the
    batched run (sampler keyed by problem index) must reproduce the single-problem runs exactly."""
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    H = 64
    imgs = [synth_image(H, H, s) for s in (1, 2, 3)]
    snrs = [10., 20., 30.]
    etas = [1500.0, 2500.0, 3500.0]
    specs, singles = [], []
    for img, snr, eta in zip(imgs, snrs, etas):
        rs = np.random.RandomState(7)
        specs.append(csmri_host_spec(img, H, H, 1.0, snr, rng=rs))
        np.random.seed(7)
        p = CSMRI(image=img, H=H, W=H, sample_prob=1.0, snr=snr)
        assert p.M0 == specs[-1]['M0'] == H * H
        singles.append(pnp_svrg(p, TVDenoiser(), eta=eta, tt=1e9, T2=4, mini_batch_size=H * H, verbose=False,
                                converge_check=False, max_iters=12, vr_mode='paper', mb_source='device', fast=True,
                                sync_every=100))
    b = BatchedSVRG(specs, T2=4, mini_batch_size=H * H, etas=etas, seed=0)
    b.run(12)
    out = b.results()
    b.close()
    assert out['z'].shape == (3, H * H) and out['psnr'].shape == (12, 3)
    for i, s in enumerate(singles):
        assert rel_l2(out['z'][i], s['z']) < 2e-6, (i, rel_l2(out['z'][i], s['z']))
        inner = [v for k, v in enumerate(s['psnr_per_iter'][1:]) if k % 5 != 0]      # drop the per-snapshot duplicates
        assert np.allclose(out['psnr'][:, i], inner, atol=0.011)


def test_batched_minibatch_run_improves_every_problem(cuda):
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    H = 128
    specs = [csmri_host_spec(synth_image(H, H, s), H, H, a, 25., rng=np.random.RandomState(s)) for s, a in
             [(0, 0.3), (1, 0.5), (2, 0.7), (3, 0.4), (4, 0.9)]]
    etas = [min(0.15 * s['M0'], 3.0 * 500) for s in specs]
    b = BatchedSVRG(specs, T2=10, mini_batch_size=500, etas=etas, seed=3)
    b.run(100)
    out = b.results()
    b.close()
    assert np.all(np.isfinite(out['z']))
    assert np.all(out['psnr'][-1] > out['psnr'][0] + 0.5), (out['psnr'][0], out['psnr'][-1])
