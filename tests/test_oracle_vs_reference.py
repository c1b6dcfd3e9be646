"""CPU, build container only: the oracle restatement against the reference imported live from
/root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/algorithms'), reason='/root/reference not present')

SCRIPT = r'''
import numpy as np
from oracle.refshim import Reference
from oracle import problems_port as PP, algorithms_port as AP
R = Reference()
img = 'data/Set12/02.png'
def cmp(refp, portp, algo, kw, budget):
    np.random.seed(3); p = refp(); d = R.TV.TVDenoiser(); R.record(p, d)
    np.random.seed(4); out = R.run(algo, p, d, budget, **kw)
    np.random.seed(3); q = portp()
    assert np.array_equal(p.Xinit, q.Xinit) and np.array_equal(p.Y, q.Y)
    np.random.seed(4); o2 = getattr(AP, algo)(q, AP.TVPort(), budget=budget, **kw)
    assert np.array_equal(out['z'], o2['z']), (algo, np.abs(out['z'] - o2['z']).max())
    assert list(out['psnr_per_iter']) == list(o2['psnr_per_iter']), algo
cs = (lambda: R.problems.CSMRI(img, H=32, W=32, sample_prob=0.4, snr=15.),
      lambda: PP.CSMRIPort(img, H=32, W=32, sample_prob=0.4, snr=15.))
for algo, kw in [('pnp_gd', dict(eta=100.)), ('pnp_sgd', dict(eta=50., mini_batch_size=60)),
                 ('pnp_svrg', dict(eta=100., T2=3, mini_batch_size=60)),
                 ('pnp_saga', dict(eta=50., mini_batch_size=60, hist_size=3)),
                 ('pnp_sarah', dict(eta=50., T2=3, mini_batch_size=60))]:
    for cc in (False, True):
        cmp(cs[0], cs[1], algo, dict(kw, converge_check=cc, diverge_check=cc), 9)
db = (lambda: R.problems.Deblur(img, H=32, W=32, kernel_path='data/kernel25.png', scale_percent=50, snr=20.),
      lambda: PP.DeblurPort(img, H=32, W=32, kernel_path='data/kernel25.png', scale_percent=50, snr=20.))
cmp(db[0], db[1], 'pnp_saga', dict(eta=0.3, mini_batch_size=20, hist_size=3, converge_check=False), 6)
pr = (lambda: R.problems.PhaseRetrieval(img, H=16, W=16, num_meas=128, snr=20.),
      lambda: PP.PhaseRetrievalPort(img, H=16, W=16, num_meas=128, snr=20.))
cmp(pr[0], pr[1], 'pnp_sarah', dict(eta=0.02, T2=3, mini_batch_size=20, converge_check=False), 6)
print('LIVE-OK')
'''


def test_port_equals_live_reference():
    # separate interpreter: the reference claims the top-level names problems/algorithms/denoisers
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-c', SCRIPT], cwd='/root/reference', env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'LIVE-OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
