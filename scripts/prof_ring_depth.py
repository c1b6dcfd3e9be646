"""Is the public graph-replay loop (mb_source='host') bound by the host or by the GPU?  Same run with staging rings
of different depth: if the time spent inside the native staging call falls with a deeper ring while the throughput
stays, the host was only waiting for the GPU (the event that guards the reuse of a staging buffer)."""
import os, sys, time, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pnp_svrg_b200 import engine as E
from pnp_svrg_b200.algorithms import pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser

iters = 2000
cfg = bench.workload(argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1))
prob = bench.make_run(cfg, seed=0)[0]
kw = dict(eta=cfg['eta'], T2=10, mini_batch_size=cfg['mini_batch_size'], vr_mode='paper', verbose=False,
          converge_check=False, mb_seed=11, fast=True, mb_source='host')
acc = {'t': 0.0}
orig = E.HostDrawRing.stage


def stage(self, *a):
    t0 = time.perf_counter()
    r = orig(self, *a)
    acc['t'] += time.perf_counter() - t0
    return r


E.HostDrawRing.stage = stage
res = {}
pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=100, **kw)
for extra in (4, 64, 4, 64):
    os.environ['PNP_HOST_RING_EXTRA'] = str(extra)
    acc['t'] = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=iters, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res.setdefault('ring_extra_%d' % extra, []).append({'it_per_s': round(iters / dt), 'stage_us_per_it': round(acc['t'] / iters * 1e6, 1)})
print(json.dumps(res))
