"""Host-drawn vs device-drawn minibatches in the whole-epoch graph (bench workload): (1) device time per inner iteration
with CUDA events over 200 epochs, (2) pure host time of one SvrgRun.epoch() call (the device drained before every call, so
nothing in it waits for the GPU), split into staging the draws and the rest.  python scripts/prof_epoch_host.py [out.json]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from pnp_svrg_b200.algorithms import SvrgRun
from pnp_svrg_b200.denoisers import TVDenoiser

a = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(a)
prob, _ = bench.make_run(cfg, seed=0)
out = {}
for src in ('host', 'device'):
    run = SvrgRun(prob, TVDenoiser(), cfg['eta'], 10, cfg['mini_batch_size'], vr_mode='paper', mb_source=src, mb_seed=11, fast=True)
    run._prepare_epochs()
    eng = run.eng
    eng.time_log.append(0.0); eng.psnr_log.append(eng.psnr_of(eng.z))
    for _ in range(30):
        run.epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(eng.stream)
    for _ in range(200):
        run.epoch()
    e1.record(eng.stream)
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    r = {'us_per_iteration_device_events': e0.elapsed_time(e1) * 1e3 / 2000, 'enqueue_ms': 1e3 * (t1 - t0), 'drain_ms': 1e3 * (t2 - t1)}
    # pure host cost of a call
    stage_t = [0.0]
    if src == 'host':
        orig = run._stage_epoch
        def timed(s_):
            t = time.time(); orig(s_); stage_t[0] += time.time() - t
        run._stage_epoch = timed
    tot = 0.0
    for _ in range(50):
        torch.cuda.synchronize()
        time.sleep(0.002)                      # the sampler threads refill their look-ahead
        t = time.time(); run.epoch(); tot += time.time() - t
    torch.cuda.synchronize()
    r['host_us_per_epoch_call_gpu_idle'] = 1e6 * tot / 50
    r['of_which_stage_us'] = 1e6 * stage_t[0] / 50
    # back to back with the GPU idle at the start: how fast can the host enqueue (no sleep)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(4):
        run.epoch()
    r['host_us_per_epoch_call_burst4'] = 1e6 * (time.time() - t) / 4
    torch.cuda.synchronize()
    run.close()
    eng.result('x')
    out[src] = r
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write(s + '\n')
