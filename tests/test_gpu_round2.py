"""GPU parity cases added in round 2 ("what's weak" 1c-1e, "what's missing" 6-7):
random-row masks, the batched sweep engine against the single engine on a SHARED minibatch stream, the tensor-core
CNN prox inside a PnP loop (final PSNR within 0.05 dB of the fp32 loop), MMO in bf16, the sharded snapshot gradient
with pruned passes, and the loud failures the advisor asked for."""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ random-row mask
def test_mask_type_rows_matches_oracle(cuda):
    """configs[0] says "random-row k-space mask": CSMRI(mask_type='rows') (additive mode, SURVEY 8(a')) samples whole
    ky rows.  Same gradient code; checked against the NumPy oracle with that mask injected."""
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    H = 128
    img = synth_image(H, H, 4)
    np.random.seed(0)
    dut = CSMRI(image=img, H=H, W=H, sample_prob=0.4, snr=20., mask_type='rows')
    rows = dut.mask[:, 0]
    assert np.array_equal(dut.mask, np.repeat(rows[:, None], H, axis=1)) and 0 < rows.sum() < H      # whole rows
    assert dut.M0 == int(rows.sum()) * H
    # the oracle restates the reference's Bernoulli mask only: build it, then give it the row mask and redo the
    # measurement model with the DUT's noise realisation (Y on the mask is what both sides must agree on)
    np.random.seed(0)
    ref = CSMRIPort(img, H=H, W=H, sample_prob=0.4, snr=20.)
    ref.mask = dut.mask.copy()
    ref.Y0, ref.Y, ref.sigma = dut.Y0.copy(), dut.Y.copy(), dut.sigma
    ref.M0 = dut.M0
    ref.Xinit = dut.Xinit.copy()
    z = np.random.default_rng(1).uniform(0, 1, H * H)
    assert rel_l2(dut.grad_full(z), ref.grad_full(z)) < 2e-6
    np.random.seed(3)
    mb_ref = ref.select_mb(700)
    np.random.seed(3)
    mb = dut.select_mb(700)
    assert np.array_equal(np.asarray(mb), mb_ref)
    assert np.all(rows[np.nonzero(mb_ref)[0]] == 1)                      # drawn from sampled rows only
    assert rel_l2(dut.grad_stoch(z, mb), ref.grad_stoch(z, mb_ref)) < 2e-6
    kw = dict(eta=0.15 * dut.M0, T2=5, mini_batch_size=700, vr_mode='paper', converge_check=False)
    np.random.seed(5)
    want = AP.pnp_svrg(ref, AP.TVPort(), budget=15, **kw)
    np.random.seed(5)
    got = pnp_svrg(dut, TVDenoiser(), tt=1e9, max_iters=15, verbose=False, **kw)
    assert rel_l2(got['z'], want['z']) < 1e-4
    assert abs(got['psnr_per_iter'][-1] - want['psnr_per_iter'][-1]) <= 0.05


# ------------------------------------------------------------------------------------------ batched vs single engine
def test_batched_engine_equals_single_engine_on_the_same_minibatches(cuda):
    """The batched sweep engine draws problem i's minibatch of iteration c with the device sampler keyed by
    (seed, c, i).  The host twin of the sampler reproduces those draws, so the single-problem engine fed the same
    positions as a stream must give the same iterates -- a real minibatch run (B << M0), not the full-support trick."""
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.engine import feistel_sample
    from pnp_svrg_b200.problems import CSMRI
    H, B, T2, n, seed = 64, 150, 4, 12, 9
    cases = [(1, 0.4, 15.), (2, 0.6, 25.), (3, 0.8, 30.)]
    specs, singles = [], []
    etas = [400.0, 600.0, 800.0]
    for (s, alpha, snr) in cases:
        specs.append(csmri_host_spec(synth_image(H, H, s), H, H, alpha, snr, rng=np.random.RandomState(s)))
    b = BatchedSVRG(specs, T2=T2, mini_batch_size=B, etas=etas, seed=seed)
    b.run(n)
    out = b.results()
    b.close()
    for i, (s, alpha, snr) in enumerate(cases):
        np.random.seed(s)
        p = CSMRI(image=synth_image(H, H, s), H=H, W=H, sample_prob=alpha, snr=snr)
        assert p.M0 == specs[i]['M0'] and np.array_equal(np.flatnonzero(p.mask), specs[i]['support'])
        sup = np.flatnonzero(p.mask)
        stream = [sup[feistel_sample(p.M0, B, seed, c, img=i)] for c in range(n)]
        one = pnp_svrg(p, TVDenoiser(), eta=etas[i], tt=1e9, T2=T2, mini_batch_size=B, verbose=False, converge_check=False,
                       max_iters=n, vr_mode='paper', mb_source='stream', mb_stream=stream, fast=True)
        assert rel_l2(out['z'][i], one['z']) < 5e-6, (i, rel_l2(out['z'][i], one['z']))
        inner = [v for k, v in enumerate(one['psnr_per_iter'][1:]) if k % (T2 + 1) != 0]
        assert np.allclose(out['psnr'][:, i], inner, atol=0.011)


# ------------------------------------------------------------------------------------------ tensor-core prox in the loop
def test_pnp_svrg_bf16_dncnn_prox_final_psnr(cuda):
    """north star: "final PSNR within 0.05 dB" for the fast mode.  PnP-SVRG (config 3 shape: T2 = 8, lr_decay 0.99) at
    256x256 with the reference's DnCNN_noise15 weights: tensor-core (bf16 operands, fp32 accumulation) prox vs the
    fp32 CUDA-core prox on the same minibatches."""
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    from pnp_svrg_b200.problems import CSMRI
    from test_gpu_cnn import _fixture
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    H = 256
    np.random.seed(0)
    p = CSMRI(image=synth_image(H, H, 0), H=H, W=H, sample_prob=0.5, snr=20.)
    kw = dict(eta=0.15 * p.M0, tt=1e9, T2=8, mini_batch_size=800, lr_decay=0.99, max_iters=48, vr_mode='paper', converge_check=False,
              verbose=False, mb_source='host', mb_seed=2)
    a = pnp_svrg(p, RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd), **kw)
    b = pnp_svrg(p, RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16'), **kw)
    assert a['psnr_per_iter'][-1] > a['psnr_per_iter'][0] + 1.0                  # the loop reconstructs
    assert abs(a['psnr_per_iter'][-1] - b['psnr_per_iter'][-1]) <= 0.05, (a['psnr_per_iter'][-1], b['psnr_per_iter'][-1])
    assert np.max(np.abs(np.array(a['psnr_per_iter']) - np.array(b['psnr_per_iter']))) <= 0.1
    assert rel_l2(b['z'], a['z']) < 5e-3


def test_mmo_bf16_matches_golden(cuda):
    """MMODenoiser(precision='bf16') (20-layer DnCNN_nobn, bias + LeakyReLU) against the reference's own output."""
    from pnp_svrg_b200.denoisers.MMODenoise import MMODenoiser
    from pnp_svrg_b200.denoisers.models.basic_models import simple_CNN
    from test_gpu_cnn import _fixture
    meta, d, sd = _fixture('ref_cnn_mmo_nobn.npz')
    mod = simple_CNN(n_ch_in=1, n_ch_out=1, n_ch=64, nl_type='relu', depth=meta['depth'], bn=False)
    mod.load_state_dict(sd)
    got = MMODenoiser(model=mod, channels=1, precision='bf16').denoise(d['noisy'])
    assert got.min() >= 0.0 and got.max() <= 1.0
    assert rel_l2(got, d['denoised']) < 5e-3, rel_l2(got, d['denoised'])
    mse = np.mean((got - d['denoised']) ** 2)
    assert 10 * np.log10(1.0 / mse) > 40.0


# ------------------------------------------------------------------------------------------ sharded snapshot
@pytest.mark.parametrize('H,world', [(128, 2), (256, 4), (512, 8), (64, 3)])
def test_band_sharded_partial_gradients(cuda, H, world):
    """config 5: every rank's partial gradient only touches its band of packed ky rows (pruned passes), the bands'
    partial gradients sum to the unsharded gradient, and the shards partition the measurements."""
    from pnp_svrg_b200.problems import CSMRI
    img = synth_image(H, H, 1)
    z = np.random.default_rng(0).uniform(0, 1, H * H)
    np.random.seed(0)
    full = CSMRI(image=img, H=H, W=H, sample_prob=0.3, snr=20.)
    g = full.grad_full(z)
    acc = np.zeros_like(g)
    cnt = 0
    for r in range(world):
        np.random.seed(0)
        p = CSMRI(image=img, H=H, W=H, sample_prob=0.3, snr=20., shard=(r, world))
        assert np.array_equal(p.mask, full.mask) and p.M0 == full.M0            # replicas of the same problem
        acc += p.grad_full(z)                                                    # no process group: the partial gradient
        cnt += p._shard_count
    assert cnt == full.M0
    assert rel_l2(acc, g) < 2e-6


def test_sharded_problem_refuses_unreduced_full_gradients(cuda):
    """pnp_gd / pnp_sarah on a sharded problem would silently use a partial gradient -> must raise."""
    from pnp_svrg_b200.algorithms import pnp_gd, pnp_sarah
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    np.random.seed(0)
    p = CSMRI(image=synth_image(64, 64, 1), H=64, W=64, sample_prob=0.3, snr=20., shard=(0, 2))
    with pytest.raises(NotImplementedError):
        pnp_gd(p, TVDenoiser(), eta=100.0, tt=1e9, max_iters=2, verbose=False)
    with pytest.raises(NotImplementedError):
        pnp_sarah(p, TVDenoiser(), eta=100.0, tt=1e9, T2=2, mini_batch_size=50, max_iters=2, verbose=False)


def test_two_rank_nccl_snapshot_allreduce(cuda):
    """The NCCL path itself: two processes, one GPU each, sharded snapshot + all-reduce vs the unsharded run.
    Needs two visible GPUs (the driver's scaling run exercises the same code through bench.py's `sharded` section)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29541', os.path.join(root, 'scripts', 'dist_check.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'OK' in r.stdout


# ------------------------------------------------------------------------------------------ loud failures
def test_oversize_minibatch_raises(cuda):
    """the device sampler's cycle walk is only defined inside [0, M0): B > M0 must raise on the host
    (the reference: np.random.choice(..., replace=False) raises ValueError)."""
    import torch
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    np.random.seed(0)
    p = CSMRI(image=synth_image(32, 32, 1), H=32, W=32, sample_prob=0.2, snr=20.)
    for src in ('device', 'host'):
        with pytest.raises(ValueError):
            pnp_svrg(p, TVDenoiser(), eta=10.0, tt=1e9, T2=2, mini_batch_size=p.M0 + 1, max_iters=2, vr_mode='paper',
                     verbose=False, mb_source=src)
    with pytest.raises(ValueError):
        p._dev_sample_sel(p._dev_new_sel(), p.M0 + 1, seed=1)
    torch.cuda.synchronize()


def test_wall_clock_budget_is_respected_in_deferred_fast_mode(cuda):
    """fast mode without stop rules must look at the clock at least every sync_every iterations."""
    import time
    from pnp_svrg_b200.algorithms import pnp_gd, pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    np.random.seed(0)
    p = CSMRI(image=synth_image(256, 256, 1), H=256, W=256, sample_prob=0.3, snr=20.)
    t = time.time()
    o = pnp_gd(p, TVDenoiser(), eta=0.15 * p.M0, tt=0.4, verbose=False, converge_check=False, fast=True)
    assert 0.35 < time.time() - t < 1.5, time.time() - t
    assert len(o['psnr_per_iter']) > 50
    t = time.time()
    o = pnp_svrg(p, TVDenoiser(), eta=0.15 * p.M0, tt=0.4, T2=10, mini_batch_size=1000, vr_mode='paper', verbose=False,
                 converge_check=False, fast=True, mb_source='device')
    assert 0.35 < time.time() - t < 1.5, time.time() - t


@pytest.mark.gpu
def test_host_draws_resolved_through_the_device_support(cuda):
    """mb_source='host' with the support list on the device (pnp_host_draws_set_device_support): the host draws ranks,
    the gather behind every staged copy turns them into the positions the host-gathering queue delivers -- one draw per
    call (stage, with an extra entry behind the indices) and an epoch per call (stage_many, consecutive ring rows)."""
    import torch
    from pnp_svrg_b200 import _lib, device as D
    from pnp_svrg_b200.engine import HostDrawRing, feistel_sample
    lib = _lib.load()
    dev = D.require_cuda()
    n, B, R, T2 = 19661, 1000, 16, 5
    sup = (np.arange(n, dtype=np.int64) * 3 + 7).astype(np.int32)
    sup_dev = torch.from_numpy(sup).to(dev)
    ring_t = torch.empty((R, B + 1), dtype=torch.int32).pin_memory()
    views = [ring_t[i].numpy() for i in range(R)]
    ring = HostDrawRing(lib, n, B, 9, None, views, 4, support_dev_ptr=D.ptr(sup_dev))
    st = torch.cuda.current_stream(dev).cuda_stream
    dst = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    for c in range(3):
        slot = ring.stage(D.ptr(dst), (41 + c,), st)
        torch.cuda.synchronize(dev)
        got = dst.cpu().numpy()
        assert np.array_equal(got[:B], sup[feistel_sample(n, B, 9, c)]), c
        assert got[B] == 41 + c                                   # the extra entry is not a rank: left alone
        assert np.array_equal(views[slot][:B], feistel_sample(n, B, 9, c))     # ranks in the staging buffer
    many = torch.zeros((T2, B + 1), dtype=torch.int32, device=dev)
    for e in range(4):                                            # wraps round the ring: runs of several rows and single rows
        ring.stage_many(D.ptr(many), T2, B + 1, st)
        torch.cuda.synchronize(dev)
        got = many.cpu().numpy()
        for j in range(T2):
            assert np.array_equal(got[j, :B], sup[feistel_sample(n, B, 9, 3 + e * T2 + j)]), (e, j)
    ring.close()
