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


def test_device_constructed_batch_is_consistent_and_reconstructs(cuda):
    """batched.csmri_device_batch builds the problems of a sweep batch on the GPU: check the pieces against each
    other and that the batched SVRG run improves every problem as it does with host-built specs."""
    import torch
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_device_batch
    H = 64
    imgs = [synth_image(H, H, k) for k in range(5)]
    alphas, snrs = [0.3, 0.5, 0.7, 0.4, 1.0], [10., 20., 30., 15., 25.]
    b = csmri_device_batch(imgs, alphas, snrs, H, H, seed=3)
    N, hp = H * H, H // 2
    m0 = b['m0_host']
    assert all(abs(m - a * N) < 5 * np.sqrt(N * a * (1 - a)) + 1 for m, a in zip(m0, alphas)) and m0[4] == N
    Y1 = torch.view_as_complex(b['Y1']).cpu().numpy()          # [nb][hp][W]   = Y[ky][kx], ky < hp
    Y2 = torch.view_as_complex(b['Y2']).cpu().numpy()          # conj Y[-ky][-kx]
    Y1n = torch.view_as_complex(b['Y1n']).cpu().numpy()        # Y[hp][kx]
    Y2n = torch.view_as_complex(b['Y2n']).cpu().numpy()
    sup = b['support'].cpu().numpy()
    for i in range(5):
        s = sup[i, :m0[i]]
        assert np.all(np.diff(s) > 0) and s.min() >= 0 and s.max() < N          # ascending flat indices, as flatnonzero
        mask = np.zeros(N, bool)
        mask[s] = True
        mask = mask.reshape(H, H)
        assert np.array_equal(np.abs(Y1[i]) > 0, mask[:hp]) and np.array_equal(np.abs(Y1n[i]) > 0, mask[hp])
        kx = np.arange(H)
        assert np.allclose(Y2[i][0], np.conj(Y1[i][0][(-kx) % H]))               # row 0 mirrors into itself
        assert np.allclose(Y2n[i], np.conj(Y1n[i][(-kx) % H]))                   # so does the Nyquist row
        assert np.allclose(Y2[i][1], np.conj(np.conj(Y2[i])[1]))                 # (trivial, keeps shapes honest)
    xi = b['xinit'].cpu().numpy()
    assert np.allclose(xi.min(axis=(1, 2)), 0, atol=1e-6) and np.allclose(xi.max(axis=(1, 2)), 1, atol=1e-6)
    xr = b['xrec'].cpu().numpy()
    for i in range(5):
        assert np.allclose(xr[i].T, (imgs[i] - imgs[i].min()) / (imgs[i].max() - imgs[i].min()), atol=1e-6)
    B = int(min(300, m0.min()))
    run = BatchedSVRG(b, T2=5, mini_batch_size=B, etas=[min(0.15 * m, 3.0 * B) for m in m0], seed=1, max_slots=64)
    run.run(40)
    out = run.results(with_z=False)
    run.close()
    # (the fully sampled problem starts from an almost exact Xinit; the prox can only lose there)
    assert np.all(out['psnr'][-1][:4] > out['psnr_init'][:4] + 0.5), (out['psnr_init'], out['psnr'][-1])
    assert np.all(np.isfinite(out['psnr'][-1]))
