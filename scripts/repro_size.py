import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from conftest import synth_image
from pnp_svrg_b200.problems import CSMRI
from pnp_svrg_b200.denoisers import TVDenoiser
from pnp_svrg_b200.algorithms import pnp_svrg
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
np.random.seed(0)
p = CSMRI(image=synth_image(H, H, 0), H=H, W=H, sample_prob=0.3, snr=20.)
for fast in (False, True):
    for src in ('device', 'host'):
        o = pnp_svrg(p, TVDenoiser(), eta=0.15 * p.M0, tt=1e9, T2=10, mini_batch_size=4000, max_iters=25, vr_mode='paper',
                     converge_check=False, verbose=False, mb_source=src, mb_seed=1, fast=fast)
        torch.cuda.synchronize()
        print(H, fast, src, o['psnr_per_iter'][0], o['psnr_per_iter'][-1], flush=True)
