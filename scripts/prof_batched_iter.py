"""Profile target (ncu): a few eager inner iterations of the batched three-pass engine at the sweep's shape
(120 problems of 256 x 256, B = 1000): selection, line pass, column + inverse line pass with the update, fused prox."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from conftest import synth_image
from pnp_svrg_b200.batched import BatchedSVRG, csmri_device_batch

nb = int(os.environ.get('NB', '120'))
images = [synth_image(256, 256, i % 12) for i in range(nb)]
alphas = [0.1 + 0.1 * (i % 9) for i in range(nb)]
snrs = [5.0 + 5.0 * (i % 7) for i in range(nb)]
batch = csmri_device_batch(images, alphas, snrs, 256, 256, seed=0, sync=False)
etas = torch.clamp(batch['m0'].to(torch.float32) * 0.15, max=3000.0)
run = BatchedSVRG(batch, T2=10, mini_batch_size=1000, etas=etas, seed=0, max_slots=64)
run.use_small = False
run.run(int(os.environ.get('ITERS', '12')))          # eager: one captured graph per inner iteration, snapshot every 10
run.results(with_z=False)
print('ok')
