"""e2e throughput of the public pnp_svrg API at the bench workload: eager vs graph-replay, host-drawn minibatches."""
import os, sys, time, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pnp_svrg_b200.algorithms import pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser

ba = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(ba)
prob_, run_ = bench.make_run(cfg, seed=0)
prob = prob_
res = {}
for fast in (False, True):
    for src in ('host', 'device'):
        kw = dict(eta=cfg['eta'], T2=10, mini_batch_size=cfg['mini_batch_size'], vr_mode='paper', verbose=False,
                  converge_check=False, mb_source=src, mb_seed=11, fast=fast)
        pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=40, **kw)
        torch.cuda.synchronize()
        t0 = time.time()
        out = pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=1000, **kw)
        torch.cuda.synchronize()
        res['fast=%s,mb=%s' % (fast, src)] = {'it_per_s': 1000 / (time.time() - t0), 'psnr_last': out['psnr_per_iter'][-1]}
print(json.dumps(res))
