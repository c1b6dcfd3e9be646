"""Where the host time of the public graph-replay loop goes (mb_source='host', bench workload): e2e it/s of three
runs, then one run with perf_counter accumulators around the pieces of an inner iteration on the host side."""
import os, sys, time, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pnp_svrg_b200 import engine as E
from pnp_svrg_b200.algorithms import _loops as L
from pnp_svrg_b200.algorithms import pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ba = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(ba)
prob = bench.make_run(cfg, seed=0)[0]
kw = dict(eta=cfg['eta'], T2=10, mini_batch_size=cfg['mini_batch_size'], vr_mode='paper', verbose=False,
          converge_check=False, mb_seed=11, fast=True)
res = {'iters': iters, 'cpus': os.cpu_count()}


def run(src, n=iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=n, mb_source=src, **kw)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0)


run('host', 40)
res['host_it_per_s'] = [round(run('host')) for _ in range(3)]
res['device_it_per_s'] = [round(run('device')) for _ in range(2)]

acc = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        acc[name + '_n'] = acc.get(name + '_n', 0) + 1
        return r
    return w


E.Engine.draw_host = timed('draw_host', E.Engine.draw_host)
E.Engine.replay = timed('graph_launch', E.Engine.replay)
E.Engine.resolve = timed('resolve(sync+readback)', E.Engine.resolve)
E.Engine.__init__ = timed('Engine.__init__', E.Engine.__init__)
E.Engine.result = timed('Engine.result', E.Engine.result)
E.HostDrawRing.stage = timed('ring.stage', E.HostDrawRing.stage)
orig = L._host_draw_fn
L._host_draw_fn = lambda eng, extra_fn=None: (lambda d: timed('draw_total', d) if d is not None else None)(orig(eng, extra_fn))
t0 = time.perf_counter()
r = run('host')
res['instrumented_it_per_s'] = round(r)
res['us_per_iteration'] = {k: round(v / iters * 1e6, 2) for k, v in acc.items() if not k.endswith('_n')}
res['calls'] = {k[:-2]: v for k, v in acc.items() if k.endswith('_n')}
print(json.dumps(res))
