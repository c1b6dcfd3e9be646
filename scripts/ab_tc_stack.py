"""One-launch conv stack (small images, bf16 tensor-core path): per-tile dataflow synchronisation between the layers
vs a grid barrier per layer (pnp_debug_set(1, 64)): identical outputs, forward time of DnCNN-17.  The dataflow form was
measured slower and is NOT in the tree: apply scripts/experiments/conv_stack_dataflow_sync.patch and rebuild first (without it
both arms run the grid-barrier kernel).
    python scripts/ab_tc_stack.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import synth_image
from test_gpu_cnn import _random_dncnn_sd
from pnp_svrg_b200 import _lib, device as D
from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
from pnp_svrg_b200.engine import ProxCtx
dev = D.require_cuda()
sd = _random_dncnn_sd(17, True, False, seed=1)
out = {}
for (H, W) in [(64, 64), (128, 128), (256, 256), (384, 256), (200, 320), (96, 32)]:
    z = D.to_lines(synth_image(H, W, 0).astype(np.float64) / 255, H, W, dev)
    res = {}
    for name, dbg in (('grid_barrier', 64), ('dataflow', 0), ('grid_barrier2', 64), ('dataflow2', 0)):
        _lib.load().pnp_debug_set(1, dbg)
        den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision='bf16')
        o = torch.empty_like(z)
        ctx = ProxCtx(z, o, H, W)
        for _ in range(3):
            den._dev_denoise(ctx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            den._dev_denoise(ctx)
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 50, o.clone())
    same = all(torch.equal(res['grid_barrier'][1], res[k][1]) for k in ('dataflow', 'dataflow2', 'grid_barrier2'))
    out['%dx%d' % (H, W)] = dict(ms={k: round(v[0], 4) for k, v in res.items()}, identical=bool(same),
                                 finite=bool(torch.isfinite(res['dataflow'][1]).all()))
    print('%dx%d' % (H, W), out['%dx%d' % (H, W)], flush=True)
_lib.load().pnp_debug_set(1, 0)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
