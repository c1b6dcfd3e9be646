"""cProfile of one rank's share of the 840-job sweep (world 8) through DeviceBatchPipeline; run under `taskset -c 0,1`
to see the host side as an 8-rank run on a 16-core box leaves it.   python scripts/prof_sweep_share_host.py"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import torch
from conftest import synth_image
from pnp_svrg_b200 import sweep as SW
images = {i: synth_image(256, 256, i) for i in range(12)}
jobs = SW.make_jobs(list(range(12)))
mine = SW.partition(jobs, 0, 8)
batch = 53
pipe = SW.DeviceBatchPipeline(H=256, W=256, iters=200, images=images)
for _ in range(2):
    pipe.submit(mine[:batch]); pipe.submit(mine[batch:2 * batch]); pipe.drain()
torch.cuda.synchronize()
for rep in range(3):
    pipe.build_seconds = 0.0
    t0 = time.time()
    SW.run_partitioned_batched(mine, pipe, 0, 1, batch=batch, gather=False)
    torch.cuda.synchronize()
    print('share %.2f ms, build host %.2f ms, cpus %d' % ((time.time() - t0) * 1e3, pipe.build_seconds * 1e3, len(os.sched_getaffinity(0))))
pr = cProfile.Profile(); pr.enable()
SW.run_partitioned_batched(mine, pipe, 0, 1, batch=batch, gather=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
