"""Profile target for the kernel families OTHER than the CSMRI iteration (ncu --set full): one or two launches of each
 -- NLM, Chambolle TV, SAGA table update, Deblur gradient (FFT and direct taps), dense PR GEMV pair, coded-diffraction
passes, fp32 conv stack, tensor-core conv stack (first / middle / last layer) -- at BASELINE config sizes."""
import os
import sys

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
from test_gpu_cnn import _random_dncnn_sd


def main():
    dev = D.require_cuda()
    H = 256
    img = synth_image(H, H, 0)
    which = set((sys.argv[1] if len(sys.argv) > 1 else 'nlm,tv,saga,pr,cdp,cnn').split(','))
    z256 = D.to_lines(img.astype(np.float64) / 255, H, H, dev)
    o256 = torch.empty_like(z256)
    if 'nlm' in which:                                   # config 2 prox
        nlm = NLMDenoiser()
        for _ in range(2):
            nlm._dev_denoise(ProxCtx(z256, o256, H, H, sigma_est=0.05))
    if 'tv' in which:                                    # Chambolle at the headline size
        Hb = 2048
        z = D.to_lines(synth_image(Hb, Hb, 0).astype(np.float64) / 255, Hb, Hb, dev)
        o = torch.empty_like(z)
        tv = TVDenoiser(method='chambolle', weight=0.1, n_iter=20)
        for _ in range(2):
            tv._dev_denoise(ProxCtx(z, o, Hb, Hb))
    if 'saga' in which:                                  # config 2: Deblur gradient (FFT path, kernel25-like) + SAGA update
        yy, xx = np.mgrid[0:H, 0:H]
        k = np.zeros((H, H))
        k[H // 2 - 12:H // 2 + 13, H // 2 - 12:H // 2 + 13] = np.round(255 * np.exp(-((yy[:25, :25] - 12) ** 2 + (xx[:25, :25] - 12) ** 2) / 50.0))
        np.random.seed(0)
        q = Deblur(image=img, H=H, W=H, kernel=k.astype(np.uint8), scale_percent=50, snr=20.)
        lip = (np.abs(np.fft.fft(q.B)).max() * np.sqrt(q.N)) ** 2
        pnp_saga(q, TVDenoiser(), eta=0.5 * q.M / lip, tt=1e9, mini_batch_size=100, hist_size=10, max_iters=3, converge_check=False,
                 verbose=False, mb_source='device')
        np.random.seed(0)
        q2 = Deblur(image=img, H=H, W=H, kernel='Minimal', scale_percent=100, snr=20.)          # direct 4-tap path
        pnp_saga(q2, TVDenoiser(), eta=1.0, tt=1e9, mini_batch_size=100, hist_size=10, max_iters=3, converge_check=False,
                 verbose=False, mb_source='device')
    if 'pr' in which:                                    # dense Gaussian PR (the reference's model), 128^2, alpha = 0.5: A = 1 GiB fp32... 64^2 here
        np.random.seed(0)
        p = PhaseRetrieval(image=synth_image(64, 64, 0), H=64, W=64, num_meas=2048, snr=20.)
        pnp_svrg(p, TVDenoiser(), eta=0.02, tt=1e9, T2=2, mini_batch_size=200, max_iters=4, vr_mode='paper', converge_check=False,
                 verbose=False, mb_source='device')
    if 'cdp' in which:                                   # config 3 gradient
        np.random.seed(0)
        c = PhaseRetrieval(image=img, H=H, W=H, model='cdp', n_masks=4, snr=20.)
        pnp_svrg(c, TVDenoiser(), eta=0.03 * c.N / (3 * np.mean(c.X ** 2)), tt=1e9, T2=2, mini_batch_size=800, max_iters=4, vr_mode='paper',
                 converge_check=False, verbose=False, mb_source='device')
    if 'cnn' in which:                                   # DnCNN-17: fp32 at 256^2, tensor cores at 2048^2
        sd = _random_dncnn_sd(17, True, False, seed=1)
        RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd)._dev_denoise(ProxCtx(z256, o256, H, H))
        Hb = 2048
        z = D.to_lines(synth_image(Hb, Hb, 0).astype(np.float64) / 255, Hb, Hb, dev)
        o = torch.empty_like(z)
        den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16')
        for _ in range(2):
            den._dev_denoise(ProxCtx(z, o, Hb, Hb))
    torch.cuda.synchronize()
    print('ok')


if __name__ == '__main__':
    main()
