"""Profile target: two DnCNN-17 forwards at 2048^2 on the tensor-core path, precision given on the command line
(bf16 | bf16x3).  ncu: -k regex:k_conv_tc -s 17 -c 2 captures two 64 -> 64 layers of the second forward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

from conftest import synth_image
from pnp_svrg_b200 import device as D
from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
from pnp_svrg_b200.engine import ProxCtx
from test_gpu_cnn import _random_dncnn_sd

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = D.require_cuda()
sd = _random_dncnn_sd(17, True, False, seed=1)
z = D.to_lines(synth_image(H, H, 0).astype(np.float64) / 255, H, H, dev)
o = torch.empty_like(z)
den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision=prec)
for _ in range(2):
    den._dev_denoise(ProxCtx(z, o, H, H))
torch.cuda.synchronize()
print('ok', prec, float(o.double().sum()))
