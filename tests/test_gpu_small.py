"""GPU parity of the one-cluster-per-image kernel (pnp_csmri_svrg_small, csrc/small.cuh) that runs whole PnP-SVRG runs
of 128^2 / 256^2 CSMRI images out of shared memory:

  * against the float64 oracle on pre-drawn minibatches (tests/test_gpu_epoch.py covers 256^2 through the public API;
    here also the random-row mask, B = 1 and B = M0);
  * against the three-pass path (PNP_SMALL=0) it replaces -- same minibatches, same logs;
  * batches (one cluster per problem, different masks / SNRs / steps) with the step decay inside the launch.

Tolerances: north star = per-iterate relative L2 <= 1e-4 in fp32, final PSNR within 0.05 dB; against the library's own
three-pass path the two differ by fp32 reduction order only (2e-6)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _pair(H, p=0.3, snr=20., seed=0, **kw):
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.problems import CSMRI
    img = synth_image(H, H, seed)
    np.random.seed(seed)
    ref = CSMRIPort(img, H=H, W=H, sample_prob=p, snr=snr)
    np.random.seed(seed)
    dut = CSMRI(image=img, H=H, W=H, sample_prob=p, snr=snr, **kw)
    if kw.get('mask_type', 'bernoulli') != 'bernoulli':
        # the oracle restates the reference's Bernoulli mask only: give it the DUT's mask and measurements
        # (tests/test_gpu_round2.py::test_mask_type_rows_matches_oracle checks the gradients of that pairing)
        ref.mask = dut.mask.copy()
        ref.Y0, ref.Y, ref.sigma, ref.M0 = dut.Y0.copy(), dut.Y.copy(), dut.sigma, dut.M0
        ref.Xinit = dut.Xinit.copy()
    return ref, dut


def _stream(prob, B, n, seed):
    rng = np.random.default_rng(seed)
    sup = np.flatnonzero(prob.mask)
    return [np.sort(rng.choice(sup, B, replace=False)).astype(np.int64) for _ in range(n)]


class _Replay:
    def __init__(self, ref, stream):
        self.ref, self.stream, self.pos = ref, stream, 0

    def __getattr__(self, k):
        return getattr(self.ref, k)

    def select_mb(self, size):
        idx = self.stream[self.pos]
        self.pos += 1
        mb = np.zeros(self.ref.N, dtype=int)
        mb[idx] = 1
        return mb.reshape(self.ref.H, self.ref.W)


def test_small_path_is_selected(cuda):
    from pnp_svrg_b200 import _lib
    from pnp_svrg_b200.algorithms import SvrgRun
    from pnp_svrg_b200.denoisers import NLMDenoiser, TVDenoiser
    lib = _lib.load()
    assert [lib.pnp_csmri_svrg_small_supported(h, w) for h, w in [(256, 256), (128, 128), (64, 64), (512, 512), (256, 128)]] == [1, 1, 0, 0, 0]
    _, dut = _pair(128)
    assert SvrgRun(dut, TVDenoiser(), 700.0, 4, 300, vr_mode='paper', mb_source='device', fast=True)._small_ok()
    assert not SvrgRun(dut, TVDenoiser(), 700.0, 4, 300, vr_mode='as_committed', mb_source='device', fast=True)._small_ok()
    assert not SvrgRun(dut, NLMDenoiser(), 700.0, 4, 300, vr_mode='paper', mb_source='device', fast=True)._small_ok()
    assert not SvrgRun(dut, TVDenoiser(method='chambolle'), 700.0, 4, 300, vr_mode='paper', mb_source='device', fast=True)._small_ok()
    with _env(PNP_SMALL='0'):
        assert not SvrgRun(dut, TVDenoiser(), 700.0, 4, 300, vr_mode='paper', mb_source='device', fast=True)._small_ok()
    args = _lib.SvrgSmallArgs(H=64, W=64, batch=1)
    assert lib.pnp_csmri_svrg_small(C.byref(args), None) == -4          # PNP_ERR_UNSUPPORTED, nothing launched


@pytest.mark.parametrize('H,B,T2,epochs,kw', [
    (256, 1000, 10, 2, {}),
    (128, 300, 5, 3, {}),
    (256, 1, 3, 2, {}),                                   # a single position per minibatch
    (128, 2000, 4, 2, dict(mask_type='rows')),            # random-row mask (configs[0])
    (256, None, 3, 2, {}),                                # B = M0: every minibatch is the whole support
])
def test_small_kernel_matches_oracle(cuda, H, B, T2, epochs, kw):
    from oracle import algorithms_port as AP
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    ref, dut = _pair(H, seed=H + T2, **kw)
    B = int(ref.M0) if B is None else B
    n = T2 * epochs
    stream = _stream(ref, B, n, seed=H)
    akw = dict(eta=min(0.15 * ref.M0, 3.0 * B), T2=T2, mini_batch_size=B, vr_mode='paper', converge_check=False, lr_decay=0.95)
    o_ref = AP.pnp_svrg(_Replay(ref, stream), AP.TVPort(), budget=n, **akw)
    o = pnp_svrg(dut, TVDenoiser(), tt=1e9, max_iters=n, verbose=False, mb_source='stream', mb_stream=stream, fast=True, **akw)
    assert len(o['psnr_per_iter']) == len(o_ref['psnr_per_iter'])
    assert rel_l2(o['z'], o_ref['z']) < 1e-4, rel_l2(o['z'], o_ref['z'])
    assert np.max(np.abs(np.array(o['psnr_per_iter']) - np.array(o_ref['psnr_per_iter']))) <= 0.0101
    assert abs(o['psnr_per_iter'][-1] - o_ref['psnr_per_iter'][-1]) <= 0.05


@pytest.mark.parametrize('H,src', [(256, 'device'), (128, 'host'), (256, 'stream')])
def test_small_kernel_equals_three_pass_path(cuda, H, src):
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    _, dut = _pair(H, p=0.4, snr=15.)
    B, T2, n = (1000 if H == 256 else 400), 6, 24
    kw = dict(eta=min(0.15 * dut.M0, 3.0 * B), T2=T2, mini_batch_size=B, vr_mode='paper', converge_check=False, verbose=False,
              tt=1e9, max_iters=n, fast=True, lr_decay=0.9, mb_source=src, mb_seed=11)
    if src == 'stream':
        kw['mb_stream'] = _stream(dut, B, n, seed=1)
    with _env(PNP_SMALL='0'):
        a = pnp_svrg(dut, TVDenoiser(sigma_modifier=1.2), **kw)
    b = pnp_svrg(dut, TVDenoiser(sigma_modifier=1.2), **kw)
    assert len(a['psnr_per_iter']) == len(b['psnr_per_iter']) == 1 + n + n // T2
    assert rel_l2(b['z'], a['z']) < 2e-6, rel_l2(b['z'], a['z'])
    assert np.allclose(a['psnr_per_iter'], b['psnr_per_iter'], atol=0.011)


def test_small_batches_use_the_cluster_kernel_up_to_the_device_capacity(cuda):
    from pnp_svrg_b200 import _lib
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    H = 128
    cap = _lib.load().pnp_csmri_svrg_small_capacity(H, H)
    assert 1 <= cap <= 148 // 8 and _lib.load().pnp_csmri_svrg_small_capacity(64, 64) == 0
    spec = csmri_host_spec(synth_image(H, H, 0), H, H, 0.5, 20., rng=np.random.RandomState(0))
    for nb, want in ((cap, True), (cap + 1, False)):
        b = BatchedSVRG([spec] * nb, T2=3, mini_batch_size=200, etas=[500.0] * nb, seed=1)
        assert b.use_small == want
        b.run(4)
        out = b.results()
        b.close()
        assert np.all(out['psnr'][-1] > out['psnr_init'])
        assert not np.array_equal(out['z'][0], out['z'][1])                # same problem, different sampler key (image index)


def test_small_kernel_batch_equals_three_pass_batch(cuda):
    """one cluster per problem, a whole multi-epoch run with step decay in one launch == the batched three-pass engine"""
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    H = 256
    cases = [(0, 0.3, 10.), (1, 0.5, 20.), (2, 0.7, 30.), (3, 1.0, 25.), (4, 0.35, 15.), (5, 0.6, 12.), (6, 0.45, 22.),
             (7, 0.8, 18.), (8, 0.3, 28.), (9, 0.55, 16.), (10, 0.4, 24.), (11, 0.9, 14.), (12, 0.65, 26.), (13, 0.5, 11.),
             (14, 0.75, 21.), (15, 0.3, 19.), (16, 0.85, 23.), (17, 0.4, 13.), (18, 0.6, 27.), (19, 0.7, 17.)]
    specs = [csmri_host_spec(synth_image(H, H, s), H, H, a, snr, rng=np.random.RandomState(s)) for s, a, snr in cases]
    etas = [min(0.15 * s['M0'], 3.0 * 1000) for s in specs]
    outs = []
    for small in ('0', '2'):           # 2: also for more problems than clusters fit the device (they run in waves)
        with _env(PNP_SMALL=small):
            b = BatchedSVRG(specs, T2=7, mini_batch_size=1000, etas=etas, seed=5, lr_decay=0.9, sigma_modifier=0.9)
            assert b.use_small == (small == '2')
            b.run(10)                  # 2 epochs, the second one cut short
            b.run(11)                  # continues: log slots, draw counters and the decayed step carry over
            outs.append(b.results())
            b.close()
    a, c = outs
    assert a['psnr'].shape == c['psnr'].shape == (21, len(cases))
    for i in range(len(cases)):
        assert rel_l2(c['z'][i], a['z'][i]) < 2e-6, (i, rel_l2(c['z'][i], a['z'][i]))
    assert np.allclose(a['psnr'], c['psnr'], atol=0.011)
    assert np.allclose(a['sigma_est'], c['sigma_est'], rtol=1e-5)
    gains = (c['psnr'][-1] > c['psnr_init']) | (np.array([a_ for _, a_, _ in cases]) == 1.0)     # fully sampled: Xinit is already the answer
    assert np.all(gains)
