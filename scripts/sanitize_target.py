"""Small end-to-end run for compute-sanitizer: every kernel family once at tiny sizes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np

from conftest import synth_image
from test_gpu_cnn import _random_dncnn_sd
from pnp_svrg_b200.algorithms import pnp_saga, pnp_sarah, pnp_svrg
from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
from pnp_svrg_b200.denoisers import NLMDenoiser, RealSN_DnCNNDenoiser, TVDenoiser
from pnp_svrg_b200.problems import CSMRI, Deblur, PhaseRetrieval

img = synth_image(64, 64, 0)
np.random.seed(0)
p = CSMRI(image=img, H=64, W=64, sample_prob=0.4, snr=20.)
for fast in (False, True):
    pnp_svrg(p, TVDenoiser(), eta=200., tt=1e9, T2=3, mini_batch_size=100, verbose=False, converge_check=False, max_iters=6,
             vr_mode='paper', mb_source='device', fast=fast)
pnp_sarah(p, NLMDenoiser(), eta=100., tt=1e9, T2=2, mini_batch_size=100, verbose=False, converge_check=False, max_iters=3)
d = Deblur(image=synth_image(32, 32, 1), H=32, W=32, kernel='Minimal', scale_percent=50, snr=20.)
pnp_saga(d, TVDenoiser(), eta=50., tt=1e9, mini_batch_size=20, hist_size=3, verbose=False, converge_check=False, max_iters=3)
r = PhaseRetrieval(image=synth_image(32, 32, 2), H=32, W=32, num_meas=128, snr=20.)
pnp_svrg(r, TVDenoiser(), eta=0.02, tt=1e9, T2=2, mini_batch_size=16, verbose=False, converge_check=False, max_iters=3, vr_mode='paper')
sd = _random_dncnn_sd(4, True, False, seed=1)
for prec in ('fp32', 'bf16'):
    RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision=prec).denoise(img.astype(np.float64) / 255)
specs = [csmri_host_spec(synth_image(64, 64, s), 64, 64, 0.5, 20., rng=np.random.RandomState(s)) for s in range(3)]
b = BatchedSVRG(specs, T2=3, mini_batch_size=100, etas=[200.] * 3)
b.run(6)
b.results()
b.close()
print('sanitize target finished')
# whole-run graph of the three-pass batched engine: first inner iteration of an epoch = update pass without a spectrum
b = BatchedSVRG(specs, T2=3, mini_batch_size=100, etas=[200.] * 3)
b.use_small = False
b.whole_run_graph = True
b.run(7)
b.results()
b.close()
# software grid barrier of the tail (whole-epoch graph at a size that takes the single-launch tail)
os.environ['PNP_SW_BARRIER'] = '1'
np.random.seed(1)
p5 = CSMRI(image=synth_image(512, 512, 3), H=512, W=512, sample_prob=0.3, snr=20.)
pnp_svrg(p5, TVDenoiser(), eta=6000., tt=1e9, T2=3, mini_batch_size=5000, verbose=False, converge_check=False, max_iters=6,
         vr_mode='paper', mb_source='device', fast=True)
os.environ['PNP_SW_BARRIER'] = '0'
print('sanitize target (round-2 additions) finished')
