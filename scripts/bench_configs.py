"""Throughput of BASELINE.json configs 2 and 3 (and the additive prox modes) through the public API.

  config 2   PnP-SAGA, Deblur 256x256 (25x25 Gaussian kernel image, scale 50 %), NLM prox
  config 3   PnP-SVRG, coded-diffraction phase retrieval 256x256 (4 masks), DnCNN-17 prox on tensor cores (bf16)
  tv         Chambolle TV prox alone, 2048x2048, 20 iterations

Synthetic image / weights (no network); mb_source='device' (device-drawn minibatches), fast mode (graph replay,
deferred PSNR read-back).  Prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

from conftest import synth_image
from pnp_svrg_b200 import device as D
from pnp_svrg_b200.algorithms import pnp_saga, pnp_svrg
from pnp_svrg_b200.denoisers import NLMDenoiser, RealSN_DnCNNDenoiser, TVDenoiser
from pnp_svrg_b200.engine import ProxCtx
from pnp_svrg_b200.problems import Deblur, PhaseRetrieval


def timed(fn, iters):
    fn(iters // 4 or 1)                      # warm-up (graph capture, allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn(iters)
    torch.cuda.synchronize()
    return iters / (time.perf_counter() - t0), out


def main():
    dev = D.require_cuda()
    res = {}
    H = 256
    img = synth_image(H, H, 0)
    # ---- config 3 ----
    from test_gpu_cnn import _random_dncnn_sd
    np.random.seed(0)
    p = PhaseRetrieval(image=img, H=H, W=H, model='cdp', n_masks=4, snr=20.)
    sd = _random_dncnn_sd(17, True, False, seed=1)
    last = max((k for k in sd if k.endswith('.weight') and sd[k].ndim == 4), key=lambda k: int(k.split('.')[-2]))
    sd[last] = sd[last] * 1e-3               # random weights are no denoiser: keep the residual small so the loop stays bounded
    den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16')
    eta = 0.03 * p.N / (3 * np.mean(p.X ** 2))
    run = lambda n: pnp_svrg(p, den, eta=eta, tt=1e9, T2=8, mini_batch_size=800, lr_decay=0.99, max_iters=n, vr_mode='paper',
                             converge_check=False, verbose=False, mb_source='device', fast=True)
    ips, out = timed(run, 400)
    res['config3_cdp256_svrg_dncnn17_bf16'] = {'inner_iterations_per_s': ips, 'psnr_first': out['psnr_per_iter'][0],
                                               'psnr_last': out['psnr_per_iter'][-1], 'n_masks': 4, 'T2': 8, 'B': 800}
    # ---- config 2 ----
    yy, xx = np.mgrid[0:H, 0:H]
    k = np.zeros((H, H))
    k[H // 2 - 12:H // 2 + 13, H // 2 - 12:H // 2 + 13] = np.round(255 * np.exp(-((yy[:25, :25] - 12) ** 2 + (xx[:25, :25] - 12) ** 2) / 50.0))
    np.random.seed(0)
    q = Deblur(image=img, H=H, W=H, kernel=k.astype(np.uint8), scale_percent=50, snr=20.)
    nlm = NLMDenoiser()
    lip = (np.abs(np.fft.fft(q.B)).max() * np.sqrt(q.N)) ** 2
    run2 = lambda n: pnp_saga(q, nlm, eta=0.5 * q.M / lip, tt=1e9, mini_batch_size=100, hist_size=10, max_iters=n, converge_check=False,
                              verbose=False, mb_source='device', fast=True)
    ips2, out2 = timed(run2, 200)
    res['config2_deblur256_saga_nlm'] = {'iterations_per_s': ips2, 'psnr_first': out2['psnr_per_iter'][0],
                                         'psnr_last': out2['psnr_per_iter'][-1], 'B': 100, 'hist': 10}
    # ---- Chambolle TV prox alone ----
    Hb = 2048
    z = D.to_lines(synth_image(Hb, Hb, 0).astype(np.float64) / 255, Hb, Hb, dev)
    o = torch.empty_like(z)
    tv = TVDenoiser(method='chambolle', weight=0.1, n_iter=20)
    ctx = ProxCtx(z, o, Hb, Hb)
    for _ in range(3):
        tv._dev_denoise(ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tv._dev_denoise(ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res['tv_chambolle_2048_20iter'] = {'ms': ms, 'gb_s_algorithmic_8B_per_px': 8.0 * Hb * Hb / ms / 1e6,
                                       'updates_per_px_per_s': 19.0 * Hb * Hb / ms * 1e3}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
