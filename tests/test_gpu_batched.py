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


@pytest.mark.parametrize('H', [64, 256])
def test_native_batch_constructor_against_numpy(cuda, H):
    """pnp_csmri_build_batch (the package's own device constructor: counter-based RNG, own FFT passes, compaction) against a
    float64 NumPy recomputation of problems/CSMRI.py:12-41 from the mask and the noise it drew: the measurements are
    mask o (fft2(X) + real noise) with the noise level of problems/problem.py:58-61, the mirrored planes, the packed
    selection bytes, M0 / 1 / M0 / the support lists and Xinit = minmax(|ifft2(Y)|) all agree; the torch.fft constructor
    of round 1 (native=False) stays available and gives the same statistics."""
    import torch
    from pnp_svrg_b200.batched import csmri_device_batch
    imgs = [synth_image(H, H, k) for k in range(3)]
    alphas, snrs = [0.3, 0.6, 1.0], [10., 20., 30.]
    b = csmri_device_batch(imgs, alphas, snrs, H, H, seed=5, native=True)
    assert b.get('native') is True
    N, hp = H * H, H // 2
    m0 = b['m0_host']
    Y1 = torch.view_as_complex(b['Y1']).cpu().numpy().astype(np.complex128)
    Y2 = torch.view_as_complex(b['Y2']).cpu().numpy().astype(np.complex128)
    Y1n = torch.view_as_complex(b['Y1n']).cpu().numpy().astype(np.complex128)
    Y2n = torch.view_as_complex(b['Y2n']).cpu().numpy().astype(np.complex128)
    bits = b['bits_full'].cpu().numpy()
    sup = b['support'].cpu().numpy()
    inv_m0 = b['inv_m0'].cpu().numpy()
    xinit = b['xinit'].cpu().numpy()
    kk = (-np.arange(H)) % H
    for i in range(3):
        x = b['xrec'][i].cpu().numpy().T.astype(np.float64)
        assert np.allclose(x, (imgs[i] - imgs[i].min()) / (imgs[i].max() - imgs[i].min()), atol=1e-6)
        s = sup[i, :m0[i]]
        assert np.all(np.diff(s) > 0)
        mask = np.zeros(N, bool)
        mask[s] = True
        mask = mask.reshape(H, H)
        assert abs(m0[i] - alphas[i] * N) < 5 * np.sqrt(N * alphas[i] * (1 - alphas[i])) + 1
        assert abs(inv_m0[i] * m0[i] - 1) < 1e-6
        # packed selection bytes of the full mask (csrc/csmri.cuh::set_sel_bits layout)
        mir = mask[kk][:, kk]
        want_bits = mask[:hp].astype(np.uint8) + 2 * mir[:hp].astype(np.uint8)
        want_bits[0] += 4 * mask[hp].astype(np.uint8) + 8 * mir[hp].astype(np.uint8)
        assert np.array_equal(bits[i], want_bits)
        # the full measurement plane from the two packed halves: rows < hp and hp from Y1 / Y1n, rows > hp from conj(Y2)
        Y = np.zeros((H, H), np.complex128)
        Y[:hp] = Y1[i]
        Y[hp] = Y1n[i]
        Ymir = np.conj(Y2[i])                                      # = Ym[-ky][-kx] for ky < hp
        for ky in range(1, hp):
            Y[H - ky] = Ymir[ky][kk]
        assert np.allclose(np.conj(Y2[i][0]), Y[0][kk]) and np.allclose(np.conj(Y2n[i]), Y[hp][kk])
        assert np.array_equal(np.abs(Y) > 0, mask)
        F = np.fft.fft2(x)
        noise = (Y - F)[mask]
        sigma = np.sqrt(np.linalg.norm((mask * F).ravel()) / 10 ** (snrs[i] / 10) / H / H)
        assert abs(b['sigma'][i] / sigma - 1) < 1e-4, (b['sigma'][i], sigma)
        scale = np.abs(F).max()
        assert np.abs(noise.imag).max() < 2e-6 * scale             # the noise is real (CSMRI.py:32-33); fp32 transform error
        assert abs(noise.real.std() / sigma - 1) < 0.05 and abs(noise.real.mean()) < 0.05 * sigma
        x0 = np.abs(np.fft.ifft2(Y))
        want = (x0 - x0.min()) / (x0.max() - x0.min())
        assert rel_l2(xinit[i].T, want) < 2e-5, rel_l2(xinit[i].T, want)
    # no read-back variant: same device results, nothing on the host
    b2 = csmri_device_batch(imgs, alphas, snrs, H, H, seed=5, native=True, sync=False)
    assert b2['m0_host'] is None and torch.equal(b2['m0'], b['m0']) and torch.equal(b2['Y1'], b['Y1'])
    assert torch.equal(b2['xinit'], b['xinit']) and torch.equal(b2['support'][0, :m0[0]], b['support'][0, :m0[0]])
    # another seed draws another mask
    b3 = csmri_device_batch(imgs, alphas, snrs, H, H, seed=6, native=True)
    assert not torch.equal(b3['bits_full'][0], b['bits_full'][0])
    # (fully sampled problem: the noise level does not depend on the mask draw, e.g. on whether it holds the DC coefficient)
    t = csmri_device_batch(imgs, alphas, snrs, H, H, seed=5, native=False)
    assert abs(t['sigma'][2] / b['sigma'][2] - 1) < 1e-4


def test_pipeline_builds_groups_inside_the_engines(cuda):
    """sweep.DeviceBatchPipeline: after the two engines exist, every further group is built straight into the idle engine's
    buffers (BatchedSVRG.build_from_images, no read-back) while the other engine runs.  Every job comes back once, with the
    PSNR gains of a working reconstruction, short tail groups are padded, and a group rebuilt with the same seed reproduces
    its problems exactly (the constructor is counter based)."""
    from pnp_svrg_b200 import sweep as SW
    H = 64
    images = {i: synth_image(H, H, i) for i in range(3)}
    jobs = [dict(id=n, image=n % 3, alpha=[0.3, 0.5, 0.8][n % 3], snr=[15., 25.][n % 2], algo='pnp_svrg', denoiser='TV') for n in range(22)]
    pipe = SW.DeviceBatchPipeline(H=H, W=H, iters=40, T2=5, mini_batch_size=200, images=images, seed=3)
    recs = SW.run_partitioned_batched(jobs, pipe, 0, 1, batch=4, gather=False)          # 6 groups: 4, 4, 4, 4, 4, 2 (padded)
    assert [r['id'] for r in recs] == list(range(22))
    gain = np.array([r['psnr_final'] - r['psnr_init'] for r in recs])
    assert np.all(np.isfinite(gain)) and np.mean(gain > 0.3) > 0.8, gain
    # groups 2.. were built inside the engines; group 4 again (ids 16..19, same seed) -> identical records
    again = SW.run_partitioned_batched(jobs[16:20], pipe, 0, 1, batch=4, gather=False)
    for a, b in zip(again, recs[16:20]):                 # (the sampler's draw counter keeps running: other minibatches)
        assert (a['id'], a['psnr_init']) == (b['id'], b['psnr_init']) and abs(a['psnr_final'] - b['psnr_final']) < 0.5
    pipe.close()
    # a minibatch larger than a problem's measurement count is reported when the records come back
    pipe = SW.DeviceBatchPipeline(H=H, W=H, iters=10, T2=5, mini_batch_size=3000, images=images, seed=3)
    with pytest.raises(ValueError):
        SW.run_partitioned_batched(jobs[:8], pipe, 0, 1, batch=4, gather=False)
    pipe.close()


@pytest.mark.parametrize('H', [64, 256])
def test_whole_run_graph_skipping_the_first_transforms_of_an_epoch_is_bit_identical(cuda, H):
    """Sweeps capture a whole run as one graph in which the first inner iteration of every epoch (z == w bit for bit,
    algorithms/pnp_svrg.py:53: the stochastic term is exactly zero) runs the update pass WITHOUT a spectrum instead of the
    selection and the three transform passes (pnp_csmri_grad, phases = 4, S = NULL).  Same bits as the full sequence,
    same minibatches afterwards (the draw counter advances either way)."""
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    specs = [csmri_host_spec(synth_image(H, H, s), H, H, a, snr, rng=np.random.RandomState(s)) for s, a, snr in
             [(0, 0.3, 20.), (1, 0.5, 10.), (2, 0.7, 30.), (3, 0.4, 25.), (4, 0.9, 15.)]]
    B = 200
    etas = [min(0.15 * s['M0'], 3.0 * B) for s in specs]
    outs = []
    for skip in (False, True):
        b = BatchedSVRG(specs, T2=10, mini_batch_size=B, etas=etas, seed=3)
        b.use_small = False                       # the three-pass kernels (what batches beyond the cluster kernel's capacity run)
        b.whole_run_graph = True
        b.skip_first = skip
        b.run(23)
        outs.append(b.results())
        b.close()
    assert np.all(np.isfinite(outs[0]['z']))
    assert np.array_equal(outs[0]['z'], outs[1]['z'])
    assert np.array_equal(outs[0]['psnr'], outs[1]['psnr'])
    # (the logged sigma is a sum of per-line estimates by double-precision atomics: equal up to the order of the additions)
    assert np.allclose(outs[0]['sigma_est'], outs[1]['sigma_est'], rtol=1e-12, atol=0.0)
    assert np.all(np.isfinite(outs[1]['psnr']))
