import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import synth_image
from test_gpu_cnn import _random_dncnn_sd
from pnp_svrg_b200 import device as D, _lib
from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
from pnp_svrg_b200.engine import ProxCtx
dev = D.require_cuda()
sd = _random_dncnn_sd(17, True, False, seed=1)
H = 2048
z = D.to_lines(synth_image(H, H, 0).astype(np.float64) / 255, H, H, dev); o = torch.empty_like(z)
den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16'); ctx = ProxCtx(z, o, H, H)
for dbg in (0, 2, 0, 2):
    _lib.load().pnp_debug_set(1, dbg)
    for _ in range(2): den._dev_denoise(ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); den._dev_denoise(ctx); e1.record(); torch.cuda.synchronize()
    print('dbg', dbg, 'ms', e0.elapsed_time(e1))
