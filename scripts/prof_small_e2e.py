import sys, os, cProfile, pstats, io, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from conftest import synth_image
from pnp_svrg_b200.algorithms import pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser
from pnp_svrg_b200.problems import CSMRI
np.random.seed(0)
p = CSMRI(image=synth_image(256,256,0), H=256, W=256, sample_prob=0.3, snr=20.)
kw = dict(eta=3000.0, T2=10, mini_batch_size=1000, vr_mode='paper', verbose=False, converge_check=False, mb_source='host', mb_seed=5, fast=True)
pnp_svrg(p, TVDenoiser(), tt=1e9, max_iters=200, **kw)
torch.cuda.synchronize()
t0=time.time(); pnp_svrg(p, TVDenoiser(), tt=1e9, max_iters=20000, **kw); torch.cuda.synchronize(); print('it/s', 20000/(time.time()-t0))
pr = cProfile.Profile(); pr.enable()
pnp_svrg(p, TVDenoiser(), tt=1e9, max_iters=20000, **kw); torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:3500])
