"""Where the time of one public pnp_svrg(..., mb_source='host', fast=True) call goes (bench workload):
python scripts/prof_e2e.py [iterations]"""
import sys, os, time, argparse, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from pnp_svrg_b200.algorithms import SvrgRun, pnp_svrg
from pnp_svrg_b200.denoisers import TVDenoiser

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
a = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(a)
prob, run0 = bench.make_run(cfg, seed=0)
kw = dict(eta=cfg['eta'], T2=10, mini_batch_size=cfg['mini_batch_size'], vr_mode='paper', verbose=False, converge_check=False,
          mb_seed=11, fast=True)
out = {}
for src in ('host', 'device'):
    pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=20, mb_source=src, **kw)
    torch.cuda.synchronize()
    t0 = time.time()
    run = SvrgRun(prob, TVDenoiser(), cfg['eta'], 10, cfg['mini_batch_size'], vr_mode='paper', mb_source=src, mb_seed=11, fast=True)
    torch.cuda.synchronize(); t1 = time.time()
    run._prepare_epochs()
    torch.cuda.synchronize(); t2 = time.time()
    eng = run.eng
    eng.time_log.append(0.0); eng.psnr_log.append(eng.psnr_of(eng.z))
    t3 = time.time()
    for _ in range(iters // 10):
        run.epoch()
    t4 = time.time()
    torch.cuda.synchronize(); t5 = time.time()
    run.close()
    res = eng.result('x')
    t6 = time.time()
    out[src] = {'init_ms': 1e3 * (t1 - t0), 'capture_ms': 1e3 * (t2 - t1), 'psnr0_ms': 1e3 * (t3 - t2), 'enqueue_ms': 1e3 * (t4 - t3),
                'drain_ms': 1e3 * (t5 - t4), 'result_ms': 1e3 * (t6 - t5), 'total_ms': 1e3 * (t6 - t0),
                'it_per_s_loop_only': iters / (t5 - t3), 'it_per_s_total': iters / (t6 - t0)}
    t0 = time.time()
    o = pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=iters, mb_source=src, **kw)
    torch.cuda.synchronize()
    out[src]['public_call_it_per_s'] = iters / (time.time() - t0)
print(json.dumps(out, indent=1))
