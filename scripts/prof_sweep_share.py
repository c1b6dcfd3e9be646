"""What one rank of an 8-rank sweep does, measured on ONE GPU: rank 0's share of the 840-job list (105 jobs) through
sweep.DeviceBatchPipeline in groups of 105 / 53 / 35 / 27 / 21, host wall clock around the call as bench_sections.sweep
takes it (no collective here), plus the GPU time of one build and one 200-iteration run of a group of each size.
    python scripts/prof_sweep_share.py [out.json]
"""
import sys, os, time, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import synth_image
from pnp_svrg_b200 import sweep as SW

world = int(os.environ.get('SHARE_WORLD', '8'))
images = {i: synth_image(256, 256, i) for i in range(12)}
jobs = SW.make_jobs(list(range(12)))
mine = SW.partition(jobs, 0, world)
out = {'share': len(mine), 'world': world, 'groups': {}}
for batch in (len(mine), -(-len(mine) // 2), -(-len(mine) // 3), -(-len(mine) // 4), -(-len(mine) // 5), -(-len(mine) // 7)):
    depth = int(os.environ.get('SHARE_DEPTH', '2'))
    pipe = SW.DeviceBatchPipeline(H=256, W=256, iters=200, images=images, depth=depth)
    for k in range(depth):                                     # warm-up: allocations and graphs of every engine
        pipe.submit(mine[:batch])
    pipe.drain()
    torch.cuda.synchronize()
    best = None
    for rep in range(5):
        pipe.build_seconds = 0.0
        t0 = time.time()
        recs = SW.run_partitioned_batched(mine, pipe, 0, 1, batch=batch, gather=False)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if best is None or dt < best[0]:
            best = (dt, pipe.build_seconds)
    # GPU time of one build and one run of this group size
    run = next(iter(pipe.engines.values()))
    g = mine[:batch]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    with torch.cuda.stream(run.stream):
        ev[0].record(run.stream)
    run.build_from_images([images[j['image']] for j in g], [j['alpha'] for j in g], [j['snr'] for j in g], 1, 0.15, 3000.0)
    with torch.cuda.stream(run.stream):
        ev[1].record(run.stream)
    run.run(200)
    with torch.cuda.stream(run.stream):
        ev[2].record(run.stream)
    torch.cuda.synchronize()
    t1 = time.time(); run.results(with_z=False); t_res = time.time() - t1
    out['groups'][batch] = dict(seconds=best[0], recon_per_s=len(mine) / best[0], build_host_s=best[1],
                                build_gpu_ms=ev[0].elapsed_time(ev[1]), run_gpu_ms=ev[1].elapsed_time(ev[2]),
                                results_host_ms=t_res * 1e3, use_small=bool(run.use_small), n_records=len(recs))
    print(batch, out['groups'][batch], flush=True)
    pipe.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
