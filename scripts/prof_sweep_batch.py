import sys, os, cProfile, pstats
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import synth_image
from pnp_svrg_b200 import sweep
images = {i: synth_image(256, 256, i) for i in range(12)}
jobs = sweep.make_jobs(list(range(12)))[:112]
sweep.reconstruct_batch(jobs[:28], H=256, W=256, iters=200, images=images)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for k in range(0, 112, 28): sweep.reconstruct_batch(jobs[k:k + 28], H=256, W=256, iters=200, images=images)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('cumulative').print_stats(40)
