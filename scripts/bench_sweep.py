"""Config 4: Set12-style sampling-ratio x SNR sweep (12 images x 10 ratios x 7 SNRs = 840 CSMRI 256x256
PnP-SVRG + wavelet-prox reconstructions, 200 inner iterations each, fixed hyper-parameters) partitioned
over the GPUs of one node.  Prints reconstructions/s (whole job).

    python scripts/bench_sweep.py [--jobs 840]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_sweep.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--jobs', type=int, default=840)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--batch', type=int, default=56, help='reconstructions per batched launch (0 = per-problem engine)')
    ap.add_argument('--repeat', type=int, default=1, help='run the 840-job list this many times (fresh draws), to amortise fixed costs')
    ap.add_argument('--construct', default='device', choices=['host', 'device'], help='where the problems of a batch are built')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from conftest import synth_image
    from pnp_svrg_b200 import sweep
    images = {i: synth_image(a.size, a.size, i) for i in range(12)}
    jobs = sweep.make_jobs(list(range(12)))[:a.jobs]
    jobs = [dict(j, id=j['id'] + r * len(jobs)) for r in range(a.repeat) for j in jobs]

    def runner(job):
        return sweep.reconstruct(job, H=a.size, W=a.size, iters=a.iters, images=images)
    runner(dict(jobs[0]))                       # warm-up (module loads, first graph)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    def batch_runner(group):
        return sweep.reconstruct_batch(group, H=a.size, W=a.size, iters=a.iters, images=images, construct=a.construct)
    if a.batch > 0:
        batch_runner(jobs[:min(a.batch, len(jobs))])
        torch.cuda.synchronize()
    t0 = time.time()
    if a.batch > 0:
        recs = sweep.run_partitioned_batched(jobs, batch_runner, rank, world, batch=a.batch, gather=True)
    else:
        recs = sweep.run_partitioned(jobs, runner, rank, world, gather=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.time() - t0
    if rank == 0:
        bad = [r for r in recs if 'error' in r]
        gain = float(np.mean([r['psnr_final'] - r['psnr_init'] for r in recs if 'error' not in r]))
        print(json.dumps({'metric': 'set12_sweep_reconstructions_per_s', 'value': len(recs) / dt, 'unit': 'recon/s',
                          'n_gpus': world, 'jobs': len(recs), 'failed': len(bad), 'seconds': dt, 'iters_per_recon': a.iters,
                          'mean_psnr_gain_db': gain, 'size': a.size, 'batch': a.batch, 'construct': a.construct,
                          'note': 'includes the problem construction (mask, fft2 measurements, Xinit) of every job, on the %s' % a.construct}))
        if bad:
            print(bad[0], file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
