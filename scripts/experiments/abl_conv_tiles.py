"""A/B of the conv tile geometry: two-dimensional tiles (default) vs one-line tiles (pnp_debug_set(1, 16)), DnCNN-17 bf16
forward, L2 flushed; python scripts/abl_conv_tiles.py [size]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import synth_image
from test_gpu_cnn import _random_dncnn_sd
from pnp_svrg_b200 import device as D, _lib
from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
from pnp_svrg_b200.engine import ProxCtx
dev = D.require_cuda()
sd = _random_dncnn_sd(17, True, False, seed=1)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
z = D.to_lines(synth_image(H, H, 0).astype(np.float64) / 255, H, H, dev); o = torch.empty_like(z)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res, outs = {}, {}
den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16'); ctx = ProxCtx(z, o, H, H)
for dbg in (0, 16, 0, 16):
    _lib.load().pnp_debug_set(1, dbg)
    for _ in range(2): den._dev_denoise(ctx)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); den._dev_denoise(ctx); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    outs[dbg] = o.clone()
    res.setdefault('one_line_tiles_ms' if dbg else 'two_dimensional_tiles_ms', []).append(float(np.median(ts)))
res['identical'] = bool(torch.equal(outs[0], outs[16]))
res['size'] = H
_lib.load().pnp_debug_set(1, 0)
print(json.dumps(res))
