import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import synth_image
from pnp_svrg_b200 import sweep
images = {i: synth_image(256, 256, i) for i in range(12)}
jobs = sweep.make_jobs(list(range(12)))
pick = [jobs[i] for i in (0, 6, 35, 69, 100, 333, 500, 839)]
a = [sweep.reconstruct(dict(j), images=images) for j in pick]
b = sweep.reconstruct_batch(pick, images=images)
for j, x, y in zip(pick, a, b):
    print(j['alpha'], j['snr'], 'single init/final %.2f %.2f' % (x['psnr_init'], x['psnr_final']), ' batched first/final %.2f %.2f' % (y['psnr_init'], y['psnr_final']))
