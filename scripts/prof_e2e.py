import sys, os, cProfile, pstats, argparse
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
a = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(a)
from pnp_svrg_b200.algorithms import pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser
from pnp_svrg_b200.problems import CSMRI
np.random.seed(0)
prob = CSMRI(image=bench.make_image(2048, 0), H=2048, W=2048, sample_prob=0.3, snr=20.)
kw = dict(eta=cfg['eta'], T2=10, mini_batch_size=cfg['mini_batch_size'], vr_mode='paper', verbose=False, converge_check=False, mb_source='host', mb_seed=11)
pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=20, **kw)
pr = cProfile.Profile(); pr.enable()
pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=200, **kw)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
