"""GPU parity: CSMRI gradients, prox kernels and the five loops against the float64 oracle.

Tolerances: north star = per-iterate relative L2 <= 1e-4 in fp32, final PSNR within 0.05 dB.
"""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu

TOL_ITER = 1e-4


def _pair(H, p=0.3, snr=20., seed=0, img_seed=0):
    from oracle.problems_port import CSMRIPort
    from pnp_svrg_b200.problems import CSMRI
    img = synth_image(H, H, img_seed)
    np.random.seed(seed)
    ref = CSMRIPort(img, H=H, W=H, sample_prob=p, snr=snr)
    np.random.seed(seed)
    dut = CSMRI(image=img, H=H, W=H, sample_prob=p, snr=snr)
    return ref, dut


@pytest.mark.parametrize('H', [32, 64, 128, 256, 512, 1024, 2048])
def test_constructor_and_grad_full(cuda, H):
    ref, dut = _pair(H)
    assert np.array_equal(ref.mask, dut.mask)
    assert dut.M0 == ref.M0 and dut.M == ref.M
    assert rel_l2(dut.Y, ref.Y) < 1e-7            # fft2 vs the reference's dense DFT product
    assert rel_l2(dut.Xinit, ref.Xinit) < 1e-7
    assert abs(dut.sigma - ref.sigma) < 1e-9 * ref.sigma
    z = np.random.default_rng(1).uniform(0, 1, ref.N)
    g_ref = ref.grad_full(z)
    g = dut.grad_full(z)
    assert g.shape == g_ref.shape and g.dtype == np.float64
    assert rel_l2(g, g_ref) < 2e-6, rel_l2(g, g_ref)


@pytest.mark.parametrize('H,B', [(32, 1), (32, 50), (64, 100), (128, 1000), (256, 1000), (512, 5000), (2048, 100000)])
def test_grad_stoch(cuda, H, B):
    ref, dut = _pair(H)
    z = np.random.default_rng(2).uniform(0, 1, ref.N)
    np.random.seed(5)
    mb_ref = ref.select_mb(B)
    np.random.seed(5)
    mb = dut.select_mb(B)
    assert np.array_equal(np.asarray(mb), mb_ref) and mb.shape == (H, H)
    g_ref = ref.grad_stoch(z, mb_ref)
    assert rel_l2(dut.grad_stoch(z, mb), g_ref) < 2e-6
    # a plain 0/1 array (no cached positions) must give the same answer
    assert rel_l2(dut.grad_stoch(z, np.array(mb_ref)), g_ref) < 2e-6


def test_grad_stoch_special_frequencies(cuda):
    """DC, Nyquist row/column and self-conjugate corners exercise the packed-column path."""
    H = 64
    ref, dut = _pair(H, p=1.0)
    z = np.random.default_rng(3).uniform(0, 1, ref.N)
    for pts in ([(0, 0)], [(H // 2, 0)], [(0, H // 2)], [(H // 2, H // 2)], [(H // 2, 5)], [(H // 2, H - 5)],
                [(0, 7)], [(0, H - 7)], [(3, 0)], [(H - 3, 0)], [(3, H // 2)], [(5, 9), (H - 5, H - 9)],
                [(1, 1)], [(H - 1, H - 1)], [(H // 2 + 1, 3)], [(H // 2 - 1, H - 3)]):
        mb = np.zeros((H, H), dtype=int)
        for (r, c) in pts:
            mb[r, c] = 1
        g_ref = ref.grad_stoch(z, mb)
        err = np.linalg.norm(dut.grad_stoch(z, mb) - g_ref) / np.linalg.norm(g_ref)
        # a single coefficient is a difference of O(N) sums: fp32 cancellation, logic errors are O(1)
        assert err < 3e-4, (pts, err)


def test_full_support_identity(cuda):
    """grad_full == grad_stoch(full-support minibatch) / M0."""
    ref, dut = _pair(64)
    z = np.random.default_rng(4).uniform(0, 1, dut.N)
    assert rel_l2(dut.grad_stoch(z, dut.mask) / dut.M0, dut.grad_full(z)) < 1e-6


@pytest.mark.parametrize('H', [32, 64, 256, 1024, 2048, 4096])
def test_estimate_sigma(cuda, H):
    import torch
    from oracle.skimage_port import estimate_sigma
    from pnp_svrg_b200 import _lib, device as D
    rng = np.random.default_rng(H)
    W = 64 if H >= 1024 else H
    z0 = synth_image(H, W, 3).astype(np.float64) / 255 + 0.05 * rng.standard_normal((H, W))
    z0[:, 1] = 0.0                      # an all-zero column: NaN estimate in skimage -> NaN mean
    z0[:8, 2] = 0.0                     # a run of exact zeros: those coefficients are masked out
    dev = D.require_cuda()
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    zl = D.to_lines(z0, H, W, dev)
    _lib.check(_lib.load().pnp_estimate_sigma(D.ptr(zl), H, W, 1, D.ptr(out), None, D.stream()))
    got = float(out[0].item()) / W
    assert np.isnan(got) and np.isnan(estimate_sigma(z0, multichannel=True, average_sigmas=True))
    z0[:, 1] = rng.standard_normal(H)
    zl = D.to_lines(z0, H, W, dev)
    out.zero_()
    _lib.check(_lib.load().pnp_estimate_sigma(D.ptr(zl), H, W, 1, D.ptr(out), None, D.stream()))
    got = float(out[0].item()) / W
    want = estimate_sigma(z0.astype(np.float32).astype(np.float64), multichannel=True, average_sigmas=True)
    assert abs(got - want) < 2e-6 * want, (got, want)


@pytest.mark.parametrize('H', [32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_wavelet_denoise(cuda, H):
    from oracle.skimage_port import denoise_wavelet
    from pnp_svrg_b200.denoisers import TVDenoiser
    rng = np.random.default_rng(H + 1)
    W = 32 if H >= 1024 else H
    z0 = synth_image(H, W, 5).astype(np.float64) / 255 + 0.05 * rng.standard_normal((H, W))
    z0 = z0.astype(np.float32).astype(np.float64)
    d = TVDenoiser()
    for s in (0.05, 0.2, 1e-3):
        want = denoise_wavelet(z0, method='BayesShrink', sigma=s, multichannel=True, rescale_sigma=True)
        got = d.denoise(z0, sigma_est=s)
        assert got.shape == (H, W)
        assert rel_l2(got, want) < 5e-6, (s, rel_l2(got, want))
    # sigma_est <= 0 -> denoise_strength * decay**t = 0 -> identity (TV.py:26)
    assert rel_l2(d.denoise(z0, sigma_est=0), z0) < 1e-6
    assert d.t == 4


def test_psnr(cuda):
    ref, dut = _pair(128)
    for seed in range(3):
        z = np.random.default_rng(seed).uniform(0, 1, ref.N)
        assert dut.PSNR(z) == ref.PSNR(z)
    assert dut.PSNR(dut.Xinit) == ref.PSNR(ref.Xinit)


def _run_both(algo, H, kw, budget, denoiser='tv', B=None, vr_mode=None, seed=1, ref_dut=None):
    from oracle import algorithms_port as AP
    from pnp_svrg_b200 import algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    ref, dut = ref_dut or _pair(H)
    extra = {} if vr_mode is None else {'vr_mode': vr_mode}
    trace = []
    np.random.seed(seed)
    o_ref = getattr(AP, algo)(ref, AP.TVPort(), budget=budget, trace=trace, converge_check=False, **kw, **extra)
    np.random.seed(seed)
    o = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=budget, verbose=False, converge_check=False,
                           **kw, **extra)
    return o_ref, o, trace


ALGOS = [
    ('pnp_gd', dict(eta=400.0)),
    ('pnp_sgd', dict(eta=150.0, mini_batch_size=200)),
    ('pnp_svrg', dict(eta=150.0, T2=4, mini_batch_size=200)),
    ('pnp_saga', dict(eta=100.0, mini_batch_size=200, hist_size=5)),
    ('pnp_sarah', dict(eta=100.0, T2=4, mini_batch_size=200)),
]


@pytest.mark.parametrize('algo,kw', ALGOS)
def test_loop_parity_faithful(cuda, algo, kw):
    o_ref, o, _ = _run_both(algo, 64, kw, budget=12, vr_mode='paper' if algo == 'pnp_svrg' else None)
    assert o['algo_name'] == o_ref['algo_name']
    assert len(o['psnr_per_iter']) == len(o_ref['psnr_per_iter']) == len(o['time_per_iter'])
    assert rel_l2(o['z'], o_ref['z']) < TOL_ITER, rel_l2(o['z'], o_ref['z'])
    assert np.max(np.abs(np.array(o['psnr_per_iter']) - np.array(o_ref['psnr_per_iter']))) <= 0.0101
    assert abs(o['psnr_per_iter'][-1] - o_ref['psnr_per_iter'][-1]) <= 0.05


def test_svrg_as_committed(cuda):
    o_ref, o, _ = _run_both('pnp_svrg', 64, dict(eta=400.0, T2=3, mini_batch_size=100), budget=9)
    assert rel_l2(o['z'], o_ref['z']) < TOL_ITER
    assert list(np.array(o['psnr_per_iter'])) == list(np.array(o_ref['psnr_per_iter']))


def test_svrg_per_iterate_256(cuda):
    """config 1 shape: 256x256, p = 0.3, B = 1000, T2 = 10, paper-mode SVRG, every iterate checked."""
    from oracle import algorithms_port as AP
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    ref, dut = _pair(256)
    kw = dict(eta=3000.0, T2=10, mini_batch_size=1000, vr_mode='paper', converge_check=False)
    for budget in (1, 5, 20, 40):
        np.random.seed(1)
        o_ref = AP.pnp_svrg(ref, AP.TVPort(), budget=budget, **kw)
        np.random.seed(1)
        o = pnp_svrg(dut, TVDenoiser(), tt=1e9, max_iters=budget, verbose=False, **kw)
        assert rel_l2(o['z'], o_ref['z']) < TOL_ITER, (budget, rel_l2(o['z'], o_ref['z']))
        assert abs(o['psnr_per_iter'][-1] - o_ref['psnr_per_iter'][-1]) <= 0.05
    assert o_ref['psnr_per_iter'][-1] > o_ref['psnr_per_iter'][0] + 1.0     # the run actually reconstructs


@pytest.mark.parametrize('algo,kw', ALGOS)
def test_loop_fast_graph_matches_faithful(cuda, algo, kw):
    """CUDA-graph replay with a pre-drawn minibatch stream == eager faithful run, bit for bit."""
    from pnp_svrg_b200 import algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    _, dut = _pair(64)
    extra = {'vr_mode': 'paper'} if algo == 'pnp_svrg' else {}
    np.random.seed(3)
    a = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=10, verbose=False, converge_check=False, **kw, **extra)
    np.random.seed(3)
    b = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=10, verbose=False, converge_check=False, fast=True,
                           sync_every=4, **kw, **extra)
    assert len(a['psnr_per_iter']) == len(b['psnr_per_iter'])
    assert rel_l2(b['z'], a['z']) < 1e-6
    assert np.allclose(a['psnr_per_iter'], b['psnr_per_iter'], atol=0.011)


def test_device_minibatch_sampler(cuda):
    import torch
    from pnp_svrg_b200 import device as D
    _, dut = _pair(128)
    sel = dut._dev_new_sel()
    for B in (1, 17, 1000, dut.M0):
        out = torch.zeros(B, dtype=torch.int32, device=dut._device)
        cnt = torch.tensor([7], dtype=torch.int32, device=dut._device)
        dut._dev_sample_sel(sel, B, seed=123, counter=cnt, idx_out=out)
        idx = out.cpu().numpy()
        assert len(np.unique(idx)) == B                      # distinct
        assert np.all(dut.mask.ravel()[idx] == 1)            # inside the sampled support
        out2 = torch.zeros(B, dtype=torch.int32, device=dut._device)
        cnt += 1
        dut._dev_sample_sel(sel, B, seed=123, counter=cnt, idx_out=out2)
        if 1 < B < dut.M0:
            assert not np.array_equal(idx, out2.cpu().numpy())
    # the bits built by the sampler give the same gradient as the explicit-index path
    z = np.random.default_rng(0).uniform(0, 1, dut.N)
    mb = np.zeros(dut.N, dtype=int)
    mb[out2.cpu().numpy()] = 1
    g1 = dut.grad_stoch(z, mb.reshape(dut.H, dut.W))
    zl = D.to_lines(z, dut.H, dut.W, dut._device)
    g = torch.empty_like(zl)
    dut._dev_grad(zl, sel=sel, g_out=g)
    assert rel_l2(D.from_lines(g, dut.H, dut.W), g1) < 1e-6
    # single-use selection: the pass that consumes it can leave it zeroed (no memset before the next draw)
    assert int(sel.sum().item()) > 0
    dut._dev_grad(zl, sel=sel, g_out=g, clear_sel=True)
    assert rel_l2(D.from_lines(g, dut.H, dut.W), g1) < 1e-6
    assert int(sel.sum().item()) == 0
    dut._dev_sample_sel(sel, 500, seed=5, counter=cnt, clear=False)
    dut._dev_sample_sel(sel, 500, seed=5, counter=cnt, idx_out=None, clear=False)      # same draw twice: idempotent bits
    n1 = int((sel != 0).sum().item())
    dut._dev_sample_sel(sel, 500, seed=5, counter=cnt, clear=True)
    assert int((sel != 0).sum().item()) == n1


def test_stop_rules(cuda):
    """converge_check stops on equal 2-dp PSNRs; diverge_check on negative PSNR (pnp_svrg.py:88-94)."""
    from pnp_svrg_b200.algorithms import pnp_gd
    from pnp_svrg_b200.denoisers import TVDenoiser
    _, dut = _pair(64)
    o = pnp_gd(dut, TVDenoiser(), eta=1e-9, tt=1e9, max_iters=50, verbose=False, converge_check=True)
    assert len(o['psnr_per_iter']) < 10
    o = pnp_gd(dut, TVDenoiser(), eta=1e7, tt=1e9, max_iters=50, verbose=False, converge_check=False,
               diverge_check=True)
    assert o['psnr_per_iter'][-1] < 0 and len(o['psnr_per_iter']) < 51


def test_wall_clock_budget(cuda):
    import time
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    _, dut = _pair(64)
    t = time.time()
    o = pnp_svrg(dut, TVDenoiser(), eta=100.0, tt=0.3, T2=5, mini_batch_size=50, verbose=False, converge_check=False)
    assert 0.25 < time.time() - t < 3.0
    assert set(o) == {'z', 'time_per_iter', 'psnr_per_iter', 'gradient_time', 'denoise_time', 'algo_name'}
    assert o['z'].shape == (dut.N,) and o['z'].dtype == np.float64


def test_rejects_host_objects(cuda):
    from oracle import algorithms_port as AP
    from pnp_svrg_b200.algorithms import pnp_gd
    _, dut = _pair(32)
    with pytest.raises(TypeError):
        pnp_gd(dut, AP.TVPort(), eta=1.0, tt=1.0, verbose=False)


@pytest.mark.parametrize('algo,kw', ALGOS)
def test_fast_mode_deferred_log_matches(cuda, algo, kw):
    """fast mode without stop rules defers the PSNR read-back; logs and iterate must equal the
    eager run's (same minibatch stream)."""
    from pnp_svrg_b200 import algorithms as ALG
    from pnp_svrg_b200.denoisers import TVDenoiser
    _, dut = _pair(64)
    extra = {'vr_mode': 'paper'} if algo == 'pnp_svrg' else {}
    np.random.seed(5)
    a = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=23, verbose=False, converge_check=False, **kw, **extra)
    np.random.seed(5)
    b = getattr(ALG, algo)(dut, TVDenoiser(), tt=1e9, max_iters=23, verbose=False, converge_check=False, fast=True,
                           sync_every=1000, **kw, **extra)
    assert len(a['psnr_per_iter']) == len(b['psnr_per_iter']) == len(b['time_per_iter'])
    assert rel_l2(b['z'], a['z']) < 1e-6
    assert np.allclose(a['psnr_per_iter'], b['psnr_per_iter'], atol=0.011)


@pytest.mark.parametrize('H,W', [(32, 32), (64, 64), (256, 256), (1024, 64), (2048, 2048)])
def test_fused_prox_equals_two_kernel_prox(cuda, H, W):
    """pnp_prox_wavelet_fused (one cooperative launch) == pnp_estimate_sigma + pnp_wavelet_denoise."""
    import torch
    from pnp_svrg_b200 import _lib, device as D
    rng = np.random.default_rng(H + W)
    z0 = synth_image(H, W, 7).astype(np.float64) / 255 + 0.05 * rng.standard_normal((H, W))
    xr = synth_image(H, W, 7).astype(np.float64) / 255
    dev = D.require_cuda()
    lib = _lib.load()
    zl, xl = D.to_lines(z0, H, W, dev), D.to_lines(xr, H, W, dev)
    outs = []
    for fused in (False, True):
        sig = torch.zeros(4, dtype=torch.float64, device=dev)
        mse = torch.zeros(4, dtype=torch.float64, device=dev)
        slot = torch.tensor([2], dtype=torch.int32, device=dev)
        o = torch.empty_like(zl)
        if fused:
            _lib.check(lib.pnp_prox_wavelet_fused(D.ptr(zl), D.ptr(o), H, W, 1, D.ptr(sig), 1.0, 0.0, D.ptr(xl), D.ptr(mse),
                                                  D.ptr(slot), D.stream()))
        else:
            _lib.check(lib.pnp_estimate_sigma(D.ptr(zl), H, W, 1, D.ptr(sig), D.ptr(slot), D.stream()))
            _lib.check(lib.pnp_wavelet_denoise(D.ptr(zl), D.ptr(o), H, W, 1, D.ptr(sig), 0.0, 1.0, 0.0, D.ptr(xl), D.ptr(mse),
                                               D.ptr(slot), D.stream()))
        outs.append((o.cpu().numpy(), sig.cpu().numpy(), mse.cpu().numpy()))
    (a, sa, ma), (b, sb, mb) = outs
    assert sa[2] > 0 and abs(sa[2] - sb[2]) <= 1e-12 * sa[2] and sb[0] == sb[1] == sb[3] == 0
    assert np.array_equal(a, b) or rel_l2(b, a) < 1e-6
    assert abs(ma[2] - mb[2]) <= 1e-6 * ma[2]


@pytest.mark.parametrize('H,W', [(1024, 1024), (2048, 2048), (512, 512)])
def test_update_prox_single_launch_matches_separate_kernels(cuda, H, W):
    """pnp_csmri_update_prox == pnp_csmri_grad(phases=4) + pnp_prox_wavelet_fused on the same spectrum."""
    import torch
    from pnp_svrg_b200 import device as D
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.engine import ProxCtx
    from pnp_svrg_b200.problems import CSMRI
    np.random.seed(5)
    p = CSMRI(image=synth_image(H, W, 2), H=H, W=W, sample_prob=0.4, snr=20.)
    dev = p._device
    rng = np.random.default_rng(1)
    z = D.to_lines(rng.random((H, W)), H, W, dev)
    w = D.to_lines(rng.random((H, W)), H, W, dev)
    mu = D.to_lines(0.01 * rng.standard_normal((H, W)), H, W, dev)
    step = torch.tensor([0.37], dtype=torch.float32, device=dev)
    sig = torch.zeros(4, dtype=torch.float64, device=dev)
    mse = torch.zeros(4, dtype=torch.float64, device=dev)
    slot = torch.tensor([1], dtype=torch.int32, device=dev)
    gk = dict(b=w, sel=None, with_y=False, gscale=0.5, vadd=mu, step_ptr=step, z_in=z)
    # separate kernels
    z1 = torch.empty_like(z)
    p._dev_grad(z, phases=3, z_out=z1, **gk)
    p._dev_grad(z, phases=4, z_out=z1, **gk)
    den = TVDenoiser(sigma_modifier=1.3)
    o1 = torch.empty_like(z)
    assert den._dev_prox_fused(ProxCtx(z1, o1, H, W, sig_log=sig, xrec=p._xrec_dev, mse_log=mse, slot=slot))
    s1, m1 = float(sig[1]), float(mse[1])
    # one launch
    sig.zero_(); mse.zero_()
    o2 = torch.empty_like(z)
    p._dev_grad(z, phases=3, z_out=o2, **gk)
    ok = p._dev_update_prox(0.5, step, mu, z, o2, sig, 1.3, 0.0, p._xrec_dev, mse, slot)
    torch.cuda.synchronize()
    if not ok:
        pytest.skip('image does not suit the resident-line kernel')
    assert torch.equal(o1, o2)
    assert float(sig[1]) == pytest.approx(s1, rel=1e-12) and float(mse[1]) == pytest.approx(m1, rel=1e-6)


def test_svrg_fast_mode_with_single_launch_tail_matches_eager_loop(cuda):
    """1024^2 is large enough for the resident-line tail: the graph-replay loop that uses it must reproduce
    the eager loop (separate kernels, PSNR read back every iteration)."""
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    H = 1024
    np.random.seed(2)
    p = CSMRI(image=synth_image(H, H, 0), H=H, W=H, sample_prob=0.3, snr=20.)
    kw = dict(eta=0.15 * p.M0, tt=1e9, T2=3, mini_batch_size=20000, max_iters=7, vr_mode='paper', converge_check=False,
              verbose=False, mb_source='host', mb_seed=9)
    a = pnp_svrg(p, TVDenoiser(), fast=False, **kw)
    b = pnp_svrg(p, TVDenoiser(), fast=True, **kw)
    assert rel_l2(b['z'], a['z']) < 1e-6
    assert np.allclose(a['psnr_per_iter'], b['psnr_per_iter'], atol=0.011)
