"""GPU parity of the CNN denoisers (fp32 CUDA-core path): against the reference's own networks and
weights (tests/golden/ref_cnn_*.npz, made by oracle/gen_golden_cnn.py) and against a torch fp32
forward of the same architecture with random weights."""
import json
import os

import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _fixture(name):
    d = np.load(os.path.join(GOLD, name), allow_pickle=False)
    import torch
    sd = {k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith('sd__')}
    return json.loads(str(d['meta'])), d, sd


def test_dncnn17_reference_weights(cuda):
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    den = RealSN_DnCNNDenoiser(meta['model_type'], meta['sigma'], state_dict=sd)
    got = den.denoise(d['noisy'], sigma_est=0.123)          # sigma_est is ignored, as in the reference
    assert got.shape == d['noisy'].shape and got.dtype == np.float64
    assert rel_l2(got, d['denoised']) < 1e-4, rel_l2(got, d['denoised'])


def test_mmo_nobn20_reference_weights(cuda):
    from pnp_svrg_b200.denoisers.MMODenoise import MMODenoiser
    from pnp_svrg_b200.denoisers.models.basic_models import simple_CNN
    meta, d, sd = _fixture('ref_cnn_mmo_nobn.npz')
    mod = simple_CNN(n_ch_in=1, n_ch_out=1, n_ch=64, nl_type='relu', depth=meta['depth'], bn=False)
    mod.load_state_dict(sd)
    den = MMODenoiser(model=mod, channels=1)
    got = den.denoise(d['noisy'])
    assert rel_l2(got, d['denoised']) < 1e-4, rel_l2(got, d['denoised'])
    assert got.min() >= 0.0 and got.max() <= 1.0 and den.t == 1


def _random_dncnn_sd(depth, bn, realsn=False, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    sd, idx = {}, 0

    def conv(co, ci):
        return torch.randn(co, ci, 3, 3, generator=g) * (0.7 / np.sqrt(9 * ci))
    pre = 'module.dncnn.' if bn else 'dncnn.'
    for i in range(depth):
        co, ci = (64, 1) if i == 0 else ((1, 64) if i == depth - 1 else (64, 64))
        sd[pre + '%d.weight' % idx] = conv(co, ci)
        if realsn:
            sd[pre + '%d.weight_orig' % idx] = torch.randn(co, ci, 3, 3, generator=g)     # must be ignored
            sd[pre + '%d.weight_u' % idx] = torch.randn(1, co, 40, 40, generator=g)
        idx += 1
        if bn and 0 < i < depth - 1:
            sd[pre + '%d.weight' % idx] = torch.rand(64, generator=g) + 0.5
            sd[pre + '%d.bias' % idx] = torch.randn(64, generator=g) * 0.1
            sd[pre + '%d.running_mean' % idx] = torch.randn(64, generator=g) * 0.1
            sd[pre + '%d.running_var' % idx] = torch.rand(64, generator=g) + 0.5
            sd[pre + '%d.num_batches_tracked' % idx] = torch.tensor(7)
            idx += 1
        if i < depth - 1:
            idx += 1        # the ReLU slot of the Sequential
    return sd


def _torch_wrapper_forward(sd, noisy, sigma):
    """torch fp32 reference of the same op on the GPU box (wrapper arithmetic of RealSN_DnCNN.py:16-40)."""
    from oracle.algorithms_port import DnCNNPort
    sd = {(k[7:] if k.startswith('module.') else k): v.double().numpy() for k, v in sd.items() if v.ndim > 0}
    ids = sorted(int(k.split('.')[1]) for k in sd if k.endswith('.weight') and sd[k].ndim == 4)
    layers = []
    for i in ids:
        bn = None
        if 'dncnn.%d.running_var' % (i + 1) in sd:
            bn = (sd['dncnn.%d.weight' % (i + 1)], sd['dncnn.%d.bias' % (i + 1)], sd['dncnn.%d.running_mean' % (i + 1)],
                  sd['dncnn.%d.running_var' % (i + 1)])
        layers.append((sd['dncnn.%d.weight' % i], bn))
    return DnCNNPort(layers, sigma), DnCNNPort(layers, sigma).denoise(noisy)


@pytest.mark.parametrize('depth,bn,realsn,H,W', [(17, True, False, 64, 64), (4, False, True, 32, 128), (17, True, True, 128, 32),
                                                 (5, True, False, 256, 256)])
def test_dncnn_random_weights_vs_torch(cuda, depth, bn, realsn, H, W):
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    sd = _random_dncnn_sd(depth, bn, realsn, seed=depth)
    noisy = synth_image(H, W, 1).astype(np.float64) / 255 * 1.2 - 0.1
    _, want = _torch_wrapper_forward(sd, noisy, 40)
    got = RealSN_DnCNNDenoiser('RealSN_DnCNN' if realsn else 'DnCNN', 40, state_dict=sd).denoise(noisy)
    assert rel_l2(got, want) < 2e-5, rel_l2(got, want)


def test_pnp_svrg_with_dncnn_prox(cuda):
    """config 3 shape in small: PnP-SVRG with the DnCNN prox, every prox on the GPU, vs the oracle loop
    whose prox is the torch CPU forward."""
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    from pnp_svrg_b200.problems import CSMRI
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    img = synth_image(64, 64, 0)
    np.random.seed(0)
    ref = CSMRIPort(img, H=64, W=64, sample_prob=0.5, snr=20.)
    np.random.seed(0)
    dut = CSMRI(image=img, H=64, W=64, sample_prob=0.5, snr=20.)
    port, _ = _torch_wrapper_forward(sd, np.zeros((8, 8)) + np.arange(8), 15)
    kw = dict(eta=800.0, T2=4, mini_batch_size=300, vr_mode='paper', converge_check=False)
    np.random.seed(1)
    want = AP.pnp_svrg(ref, port, budget=8, **kw)
    np.random.seed(1)
    got = pnp_svrg(dut, RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd), tt=1e9, max_iters=8, verbose=False, **kw)
    assert rel_l2(got['z'], want['z']) < 1e-4, rel_l2(got['z'], want['z'])
    assert abs(got['psnr_per_iter'][-1] - want['psnr_per_iter'][-1]) <= 0.05
    assert want['psnr_per_iter'][-1] > want['psnr_per_iter'][0]


@pytest.mark.parametrize('depth,H,W', [(3, 64, 64), (3, 32, 128), (4, 128, 32), (17, 64, 64), (17, 256, 256)])
def test_tensor_core_path_matches_fp32_path(cuda, depth, H, W):
    """bf16 tcgen05 conv stack vs the fp32 CUDA-core stack: same weights, same input.  bf16 operands
    bound the agreement (~3 significant digits per layer)."""
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    sd = _random_dncnn_sd(depth, depth > 4, False, seed=100 + depth)
    noisy = synth_image(H, W, 2).astype(np.float64) / 255
    a = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd).denoise(noisy)
    b = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16').denoise(noisy)
    err = rel_l2(b, a)
    assert err < 1e-2, err


def test_tensor_core_dncnn17_reference_weights_psnr(cuda):
    """DnCNN-17 with the reference's weights on the tensor-core path: PSNR of the denoised image within
    0.05 dB of the reference's own output (north star tolerance for the fast mode)."""
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    got = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16').denoise(d['noisy'])
    clean = synth_image(64, 64, 0) * 0.0 + d['denoised']          # reference output as the yardstick
    mse = np.mean((got - clean) ** 2)
    assert 10 * np.log10(1.0 / mse) > 40.0, 10 * np.log10(1.0 / mse)     # > 40 dB agreement with the reference output
    assert rel_l2(got, d['denoised']) < 5e-3


@pytest.mark.parametrize('depth,H,W', [(3, 64, 64), (4, 32, 128), (17, 64, 64), (17, 256, 256), (17, 512, 128)])
def test_error_compensated_tensor_core_mode_matches_fp32_path(cuda, depth, H, W):
    """precision='bf16x3' (hi + lo bf16 operands, three tcgen05 products per tile into the fp32 accumulator) against the
    fp32 CUDA-core stack, same weights, same input: <= 1e-4 through 17 layers (the plain bf16 mode: ~1e-2), so the
    tensor cores also serve the exact mode (the reference runs the nets in fp32, denoisers/RealSN_DnCNN.py:32-35)."""
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    sd = _random_dncnn_sd(depth, depth > 4, False, seed=100 + depth)
    noisy = synth_image(H, W, 2).astype(np.float64) / 255
    a = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd).denoise(noisy)
    b = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16x3').denoise(noisy)
    c = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16').denoise(noisy)
    err, err_fast = rel_l2(b, a), rel_l2(c, a)
    assert err < 1e-4, (err, err_fast)
    assert err < 0.05 * err_fast, (err, err_fast)           # two orders of magnitude closer than the fast mode


def test_error_compensated_mode_reference_weights(cuda):
    """DnCNN-17 with the reference's own weights and the reference's own output (tests/golden/ref_cnn_dncnn15.npz, made
    by the unmodified torch model): the error-compensated tensor-core mode holds the tolerance of the fp32 path."""
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    got = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16x3').denoise(d['noisy'])
    fp32 = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd).denoise(d['noisy'])
    assert rel_l2(got, d['denoised']) < 1e-4, rel_l2(got, d['denoised'])
    assert rel_l2(got, fp32) < 1e-4, rel_l2(got, fp32)


def test_mmo_error_compensated_mode(cuda):
    """MMODenoiser(precision='bf16x3') (20-layer DnCNN_nobn, bias + LeakyReLU) against the reference's own output at the
    fp32 path's tolerance."""
    from pnp_svrg_b200.denoisers.MMODenoise import MMODenoiser
    from pnp_svrg_b200.denoisers.models.basic_models import simple_CNN
    meta, d, sd = _fixture('ref_cnn_mmo_nobn.npz')
    mod = simple_CNN(n_ch_in=1, n_ch_out=1, n_ch=64, nl_type='relu', depth=meta['depth'], bn=False)
    mod.load_state_dict(sd)
    got = MMODenoiser(model=mod, channels=1, precision='bf16x3').denoise(d['noisy'])
    assert got.min() >= 0.0 and got.max() <= 1.0
    assert rel_l2(got, d['denoised']) < 1e-4, rel_l2(got, d['denoised'])


def test_pnp_svrg_with_dncnn_prox_on_tensor_cores_exact_mode(cuda):
    """The loop-level parity test of test_pnp_svrg_with_dncnn_prox with the prox on the tensor cores in the
    error-compensated mode: same tolerance against the oracle loop (torch CPU fp32 forward as the prox)."""
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    from pnp_svrg_b200.problems import CSMRI
    meta, d, sd = _fixture('ref_cnn_dncnn15.npz')
    img = synth_image(64, 64, 0)
    np.random.seed(0)
    ref = CSMRIPort(img, H=64, W=64, sample_prob=0.5, snr=20.)
    np.random.seed(0)
    dut = CSMRI(image=img, H=64, W=64, sample_prob=0.5, snr=20.)
    port, _ = _torch_wrapper_forward(sd, np.zeros((8, 8)) + np.arange(8), 15)
    kw = dict(eta=800.0, T2=4, mini_batch_size=300, vr_mode='paper', converge_check=False)
    np.random.seed(1)
    want = AP.pnp_svrg(ref, port, budget=8, **kw)
    np.random.seed(1)
    got = pnp_svrg(dut, RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16x3'), tt=1e9, max_iters=8, verbose=False, **kw)
    assert rel_l2(got['z'], want['z']) < 1e-4, rel_l2(got['z'], want['z'])
    assert abs(got['psnr_per_iter'][-1] - want['psnr_per_iter'][-1]) <= 0.05
