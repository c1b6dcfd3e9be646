"""GPU parity: Deblur, PhaseRetrieval and NLM against the float64 oracle."""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


def _kernel(H, seed=0):
    rng = np.random.default_rng(seed)
    k = np.zeros((H, H))
    c = H // 2
    yy, xx = np.mgrid[0:H, 0:H]
    k += 60 * np.exp(-((yy - c) ** 2 + (xx - c - 3) ** 2) / (2 * (H / 20.0) ** 2))
    k[rng.integers(0, H, 8), rng.integers(0, H, 8)] += 20
    return np.round(k).astype(np.uint8)


def _deblur_pair(H, scale, kernel='img', seed=0):
    from oracle.problems_port import DeblurPort
    from pnp_svrg_b200.problems import Deblur
    img = synth_image(H, H, 2)
    kw = dict(H=H, W=H, scale_percent=scale, snr=20.)
    if kernel == 'img':
        kp, kd = dict(kernel_path=_kernel(H)), dict(kernel=_kernel(H))
    else:
        kp = kd = dict(kernel=kernel)
    np.random.seed(seed)
    ref = DeblurPort(img, **kw, **kp)
    np.random.seed(seed)
    dut = Deblur(image=img, **kw, **kd)
    return ref, dut


@pytest.mark.parametrize('H,scale,kernel', [(32, 50, 'img'), (64, 100, 'img'), (64, 50, 'Minimal'), (64, 25, 'Identity'),
                                            (128, 50, 'img'), (256, 50, 'img'), (256, 100, 'Minimal'), (512, 50, 'img')])
def test_deblur_constructor_and_grads(cuda, H, scale, kernel):
    ref, dut = _deblur_pair(H, scale, kernel)
    assert dut.M == ref.M and dut.lrH == ref.lrH
    assert np.array_equal(dut.B, ref.B)
    assert rel_l2(dut.Y, ref.Y) < 1e-12 and np.array_equal(dut.Xinit, ref.Xinit)
    z = np.random.default_rng(1).uniform(0, 1, ref.N)
    g_ref = ref.grad_full(z)
    assert rel_l2(dut.grad_full(z), g_ref) < 5e-6, rel_l2(dut.grad_full(z), g_ref)
    for B in (1, 37, min(1000, ref.M)):
        np.random.seed(3)
        mb_ref = ref.select_mb(B)
        np.random.seed(3)
        mb = dut.select_mb(B)
        assert np.array_equal(np.asarray(mb), mb_ref)
        g_ref = ref.grad_stoch(z, mb_ref)
        assert rel_l2(dut.grad_stoch(z, mb), g_ref) < 5e-6, (B, rel_l2(dut.grad_stoch(z, mb), g_ref))


def test_deblur_helical_boundary(cuda):
    """The raveled 1-D convolution wraps rows into each other: a tap at flat offset 1 moves the last
    pixel of a row to the first pixel of the NEXT row."""
    from pnp_svrg_b200.problems import Deblur
    H = 32
    k = np.zeros(H * H)
    k[1] = H * H / np.sqrt(H * H)          # fft_blur scales by sqrt(N) and B = kernel / N -> unit gain
    dut = Deblur(image=synth_image(H, H, 1), H=H, W=H, kernel=k, scale_percent=100, sigma=0.0)
    x = dut.X.reshape(H, H)
    y = dut.Y.reshape(H, H)
    assert np.allclose(y[3, 1:], x[3, :-1], atol=1e-9) and np.allclose(y[4, 0], x[3, -1], atol=1e-9)
    # device gradient at z = 0: B^T (0 - y) / M -> shifts back by one flat position
    g = dut.grad_full(np.zeros(H * H)).reshape(H, H)
    assert rel_l2(g, -x / dut.M) < 1e-5


def _pr_pair(H, M, seed=0):
    from oracle.problems_port import PhaseRetrievalPort
    from pnp_svrg_b200.problems import PhaseRetrieval
    img = synth_image(H, H, 4)
    np.random.seed(seed)
    ref = PhaseRetrievalPort(img, H=H, W=H, num_meas=M, snr=20.)
    np.random.seed(seed)
    dut = PhaseRetrieval(image=img, H=H, W=H, num_meas=M, snr=20.)
    return ref, dut


@pytest.mark.parametrize('H,M', [(32, 512), (32, 100), (64, 2048)])
def test_pr_constructor_and_grads(cuda, H, M):
    ref, dut = _pr_pair(H, M)
    assert np.array_equal(dut.A, ref.A) and np.array_equal(dut.Y, ref.Y)
    assert rel_l2(dut.Xinit, ref.Xinit) < 1e-6          # matrix-free power iteration vs the N x N matrix
    z = np.random.default_rng(1).uniform(0, 1, ref.N)
    assert rel_l2(dut.grad_full(z), ref.grad_full(z)) < 5e-6
    for B in (1, 33, min(800, M)):
        np.random.seed(3)
        mb_ref = ref.select_mb(B)
        np.random.seed(3)
        mb = dut.select_mb(B)
        assert np.array_equal(np.asarray(mb), mb_ref)
        assert rel_l2(dut.grad_stoch(z, mb), ref.grad_stoch(z, mb_ref)) < 5e-6


@pytest.mark.parametrize('H,W', [(32, 32), (64, 32), (128, 128)])
def test_nlm_denoise(cuda, H, W):
    from oracle.skimage_port import denoise_nl_means
    from pnp_svrg_b200.denoisers import NLMDenoiser
    rng = np.random.default_rng(H)
    clean = synth_image(H, W, 6).astype(np.float64) / 255
    for s in (0.03, 0.1):
        z0 = (clean + s * rng.standard_normal((H, W))).astype(np.float32).astype(np.float64)
        want = denoise_nl_means(z0, h=s, sigma=s, fast_mode=False, patch_size=4, patch_distance=5)
        got = NLMDenoiser().denoise(z0, sigma_est=s)
        assert got.shape == (H, W)
        bad = np.abs(got - want) > 1e-4
        # isolated pixels may flip a hard decision (cut-off 5.0) in fp32; everything else is tight
        assert bad.mean() < 2e-3, bad.mean()
        assert rel_l2(got, want) < 1e-4, rel_l2(got, want)
    d = NLMDenoiser(denoise_strength=0.05)
    want = denoise_nl_means(z0, h=0.05, sigma=0.0, fast_mode=False, patch_size=4, patch_distance=5)
    assert rel_l2(d.denoise(z0, sigma_est=0), want) < 1e-4      # sigma_est <= 0 branch (NLM.py:27)
    assert d.t == 1


@pytest.mark.parametrize('algo,kw', [('pnp_saga', dict(eta=0.3, mini_batch_size=100, hist_size=5)),
                                     ('pnp_svrg', dict(eta=0.5, T2=4, mini_batch_size=100, vr_mode='paper')),
                                     ('pnp_sarah', dict(eta=0.3, T2=3, mini_batch_size=100))])
def test_deblur_loops(cuda, algo, kw):
    from oracle import algorithms_port as AP
    from pnp_svrg_b200 import algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    ref, dut = _deblur_pair(64, 50)
    np.random.seed(1)
    o_ref = getattr(AP, algo)(ref, AP.TVPort(), budget=8, converge_check=False, **kw)
    np.random.seed(1)
    o = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=8, verbose=False, converge_check=False, **kw)
    assert rel_l2(o['z'], o_ref['z']) < 1e-4, rel_l2(o['z'], o_ref['z'])
    assert abs(o['psnr_per_iter'][-1] - o_ref['psnr_per_iter'][-1]) <= 0.05


@pytest.mark.parametrize('algo,kw', [('pnp_svrg', dict(eta=0.05, T2=4, mini_batch_size=64, vr_mode='paper')),
                                     ('pnp_sarah', dict(eta=0.03, T2=3, mini_batch_size=64)),
                                     ('pnp_gd', dict(eta=0.05))])
def test_pr_loops(cuda, algo, kw):
    from oracle import algorithms_port as AP
    from pnp_svrg_b200 import algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    ref, dut = _pr_pair(32, 512)
    dut.Xinit = ref.Xinit.copy()
    np.random.seed(1)
    o_ref = getattr(AP, algo)(ref, AP.TVPort(), budget=8, converge_check=False, **kw)
    np.random.seed(1)
    o = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=8, verbose=False, converge_check=False, **kw)
    assert rel_l2(o['z'], o_ref['z']) < 1e-4, rel_l2(o['z'], o_ref['z'])
    assert abs(o['psnr_per_iter'][-1] - o_ref['psnr_per_iter'][-1]) <= 0.05


def test_device_sampler_matches_host_twin(cuda):
    """csrc feistel_perm == engine.feistel_sample (mb_source='device' and 'host' draw identically)."""
    import torch
    from pnp_svrg_b200 import _lib, device as D
    from pnp_svrg_b200.engine import feistel_sample
    dev = D.require_cuda()
    for n, c in [(1, 1), (100, 100), (16384, 800), (1258000, 100000)]:
        out = torch.zeros(c, dtype=torch.int32, device=dev)
        cnt = torch.tensor([5], dtype=torch.int32, device=dev)
        _lib.check(_lib.load().pnp_sample_indices(D.ptr(out), n, c, 77, D.ptr(cnt), D.stream()))
        assert np.array_equal(out.cpu().numpy().astype(np.int64), feistel_sample(n, c, 77, 5))


@pytest.mark.parametrize('H,W', [(32, 32), (96, 40), (256, 256), (130, 67)])
def test_tv_chambolle(cuda, H, W):
    """Additive TVDenoiser(method='chambolle') against the skimage restatement (fixed iteration count)."""
    from oracle.skimage_port import denoise_tv_chambolle
    from pnp_svrg_b200.denoisers import TVDenoiser
    rng = np.random.default_rng(H + W)
    clean = synth_image(H, W, 3).astype(np.float64) / 255
    z0 = (clean + 0.08 * rng.standard_normal((H, W))).astype(np.float32).astype(np.float64)
    for w, n_iter in ((0.1, 1), (0.1, 2), (0.05, 7), (0.2, 13), (0.1, 40)):      # 1, 2: edge cases; 7, 13, 40: 1..7 launches
        want = denoise_tv_chambolle(z0, weight=w, eps=0.0, n_iter_max=n_iter)
        got = TVDenoiser(method='chambolle', weight=w, n_iter=n_iter).denoise(z0)
        assert got.shape == (H, W)
        assert rel_l2(got, want) < 2e-6, (w, n_iter, rel_l2(got, want))
    # weight from the loops' sigma estimate, and the weight <= 0 identity
    d = TVDenoiser(method='chambolle', n_iter=10, sigma_modifier=2.0)
    assert rel_l2(d.denoise(z0, sigma_est=0.04), denoise_tv_chambolle(z0, weight=0.08, eps=0.0, n_iter_max=10)) < 2e-6
    assert np.allclose(TVDenoiser(method='chambolle').denoise(z0, sigma_est=0), z0.astype(np.float32), atol=0)


def test_tv_chambolle_in_svrg_loop(cuda):
    """The chambolle prox inside PnP-SVRG (paper-mode VR) tracks the oracle loop with the same prox."""
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    H = 64
    img = synth_image(H, H, 1)
    np.random.seed(3)
    ref = CSMRIPort(img, H=H, W=H, sample_prob=0.5, snr=20.)
    np.random.seed(3)
    dut = CSMRI(image=img, H=H, W=H, sample_prob=0.5, snr=20.)
    kw = dict(eta=0.5 * ref.M0 / 1.0, T2=4, mini_batch_size=300)
    np.random.seed(4)
    want = AP.pnp_svrg(ref, AP.ChambollePort(n_iter=12, sigma_modifier=1.5), budget=12, vr_mode='paper', converge_check=False,
                       eta=kw['eta'], T2=4, mini_batch_size=300)
    np.random.seed(4)
    got = pnp_svrg(dut, TVDenoiser(method='chambolle', n_iter=12, sigma_modifier=1.5), tt=1e9, max_iters=12, vr_mode='paper',
                   converge_check=False, verbose=False, **kw)
    assert rel_l2(got['z'], want['z']) < 1e-4
    assert abs(got['psnr_per_iter'][-1] - want['psnr_per_iter'][-1]) <= 0.05


@pytest.mark.parametrize('H,W,L', [(32, 32, 2), (64, 32, 3), (128, 256, 1), (256, 256, 4)])
def test_cdp_constructor_and_grads(cuda, H, W, L):
    """Additive PhaseRetrieval(model='cdp') against the float64 restatement."""
    from oracle.problems_port import CDPPort
    from pnp_svrg_b200.problems import PhaseRetrieval
    img = synth_image(H, W, 4)
    np.random.seed(7)
    ref = CDPPort(img, H=H, W=W, n_masks=L, snr=20.)
    np.random.seed(7)
    dut = PhaseRetrieval(image=img, H=H, W=W, model='cdp', n_masks=L, snr=20.)
    assert dut.M == ref.M == L * H * W
    assert np.array_equal(dut.codes, ref.codes) and np.allclose(dut.Y, ref.Y) and np.allclose(dut.Xinit, ref.Xinit)
    rng = np.random.default_rng(H)
    z = rng.random(H * W)
    assert rel_l2(dut.grad_full(z), ref.grad_full(z)) < 5e-6
    np.random.seed(8)
    mb_ref = ref.select_mb(200)
    np.random.seed(8)
    mb = dut.select_mb(200)
    assert np.array_equal(np.asarray(mb), mb_ref)
    assert rel_l2(dut.grad_stoch(z, mb), ref.grad_stoch(z, mb_ref)) < 5e-6
    assert rel_l2(dut.grad_stoch(z, mb), ref.grad_stoch(z, mb_ref)) < 5e-6      # the selection scratch was left clean
    assert rel_l2(dut.grad_full(z), ref.grad_full(z)) < 5e-6


@pytest.mark.parametrize('algo,kw', [('pnp_svrg', dict(T2=4, mini_batch_size=400, vr_mode='paper')),
                                     ('pnp_sgd', dict(mini_batch_size=400)),
                                     ('pnp_sarah', dict(T2=3, mini_batch_size=400))])
def test_cdp_loops(cuda, algo, kw):
    from oracle import algorithms_port as AP
    from oracle.problems_port import CDPPort
    import pnp_svrg_b200.algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import PhaseRetrieval
    H = 32
    img = synth_image(H, H, 5)
    np.random.seed(11)
    ref = CDPPort(img, H=H, W=H, n_masks=4, snr=25.)
    np.random.seed(11)
    dut = PhaseRetrieval(image=img, H=H, W=H, model='cdp', n_masks=4, snr=25.)
    eta = 0.05 / max(np.linalg.norm(ref.X) ** 2 / ref.N, 1e-9) / 4
    okw = {k: v for k, v in kw.items()}
    np.random.seed(12)
    want = getattr(AP, algo)(ref, AP.TVPort(), eta=eta, budget=9, converge_check=False, **okw)
    np.random.seed(12)
    got = getattr(ALG, algo)(dut, TVDenoiser(), eta=eta, tt=1e9, max_iters=9, converge_check=False, verbose=False, **kw)
    assert rel_l2(got['z'], want['z']) < 1e-4, rel_l2(got['z'], want['z'])
    assert abs(got['psnr_per_iter'][-1] - want['psnr_per_iter'][-1]) <= 0.05


@pytest.mark.parametrize('kernel', ['Minimal', 'Identity', 'sparse'])
def test_deblur_direct_taps_match_fft_path(cuda, kernel):
    """Additive conv='direct' (tap sums, SURVEY 8(a')) against the four-step FFT convolution and the oracle."""
    from oracle.problems_port import DeblurPort
    from pnp_svrg_b200.problems import Deblur
    H = 64
    img = synth_image(H, H, 3)
    if kernel == 'sparse':                                  # taps that wrap around rows and around the whole image
        k = np.zeros((H, H))
        k[0, 0], k[0, H - 1], k[H - 1, H - 1], k[H - 1, 0], k[17, 5], k[40, 63] = 3, 2, 5, 1, 4, 2
        kd = kp = dict(kernel=k)
    else:
        kd = kp = dict(kernel=kernel)
    kw = dict(H=H, W=H, scale_percent=50, snr=25.)
    np.random.seed(3)
    ref = DeblurPort(img, **kw, **({'kernel_path': kp['kernel']} if kernel == 'sparse' else kp))
    np.random.seed(3)
    a = Deblur(image=img, conv='direct', **kw, **kd)
    np.random.seed(3)
    b = Deblur(image=img, conv='fft', **kw, **kd)
    assert a._direct and not b._direct
    z = np.random.default_rng(0).random(H * H)
    ga, gb, gr = a.grad_full(z), b.grad_full(z), ref.grad_full(z)
    assert rel_l2(ga, gr) < 5e-6 and rel_l2(gb, gr) < 5e-6 and rel_l2(ga, gb) < 5e-6
    np.random.seed(4)
    mb = a.select_mb(150)
    assert rel_l2(a.grad_stoch(z, mb), ref.grad_stoch(z, np.asarray(mb))) < 5e-6
    with pytest.raises(ValueError):
        Deblur(image=img, conv='direct', kernel=np.ones((H, H)), **kw)


@pytest.mark.parametrize('H,W,L', [(64, 64, 4), (128, 32, 2), (256, 256, 4)])
def test_cdp_two_points_in_one_launch_sequence(cuda, monkeypatch, H, W, L):
    """The SVRG / SARAH difference g_B(z) - g_B(w) of the coded-diffraction model (not linear: two evaluations): with the
    second scratch both points share the three passes' launches (pnp_cdp_grad, S2) -- bit-identical to the two-sequence
    form (PNP_CDP_TWO_PASS=1), for a minibatch and for the full set, and the single-use selection is left cleared."""
    import torch
    from pnp_svrg_b200 import device as D
    from pnp_svrg_b200.problems import PhaseRetrieval
    np.random.seed(2)
    p = PhaseRetrieval(image=synth_image(H, W, 1), H=H, W=W, model='cdp', n_masks=L, snr=20.)
    dev = p._device
    rng = np.random.default_rng(0)
    z = D.to_lines(p.Xinit + 0.05 * rng.standard_normal(p.N), H, W, dev)
    w = D.to_lines(p.Xinit, H, W, dev)
    mu = D.to_lines(0.01 * rng.standard_normal(p.N), H, W, dev)
    sel = torch.from_numpy(rng.choice(p.M, size=min(500, p.M // 4), replace=False).astype(np.int32)).to(dev)
    outs = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('PNP_CDP_TWO_PASS', mode)
        for name, s in (('mb', sel), ('full', None)):
            g, v, zo = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
            p._dev_grad_cdp(z, w, s, 1.0 / 500, 0.3, None, g, mu, v, z, zo)
            torch.cuda.synchronize()
            outs[(mode, name)] = (g.clone(), v.clone(), zo.clone())
            assert int(p._mask.sum()) == 0
    for name in ('mb', 'full'):
        for a, b in zip(outs[('0', name)], outs[('1', name)]):
            assert torch.equal(a, b), name
    assert float(outs[('0', 'mb')][0].abs().max()) > 0
