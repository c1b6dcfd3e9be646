"""DnCNN-17 forward timing: fp32 CUDA-core path vs the bf16 and the error-compensated bf16x3 tcgen05 paths (CUDA events, L2
flushed), and the agreement of the tensor-core paths with the fp32 path."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='256,2048')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--precisions', default='fp32,bf16')
    a = ap.parse_args()
    from conftest import synth_image
    from test_gpu_cnn import _random_dncnn_sd
    from pnp_svrg_b200 import device as D
    from pnp_svrg_b200.denoisers import RealSN_DnCNNDenoiser
    from pnp_svrg_b200.engine import ProxCtx
    dev = D.require_cuda()
    sd = _random_dncnn_sd(17, True, False, seed=1)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out, outs, errs = {}, {}, {}
    for H in [int(x) for x in a.sizes.split(',')]:
        z = D.to_lines(synth_image(H, H, 0).astype(np.float64) / 255, H, H, dev)
        o = torch.empty_like(z)
        for prec in a.precisions.split(','):
            den = RealSN_DnCNNDenoiser('DnCNN', 15, state_dict=sd, precision=prec)
            ctx = ProxCtx(z, o, H, H)
            for _ in range(2):
                den._dev_denoise(ctx)
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.iters if (prec != 'fp32' or H <= 512) else 2):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                den._dev_denoise(ctx)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            if H <= 512:                    # agreement with the fp32 CUDA-core path (same weights, same input)
                ref_out = outs.setdefault(H, o.clone()) if prec == 'fp32' else outs.get(H)
                if ref_out is not None and prec != 'fp32':
                    errs['%d_%s' % (H, prec)] = float((o - ref_out).norm() / ref_out.norm())
            flop = 1108224.0 * H * H
            mid = 15 * 2 * 9 * 64 * 64 * H * H
            out['%d_%s' % (H, prec)] = dict(ms=ms, tflops_all=flop / ms / 1e9, tflops_mid_layers_only=mid / ms / 1e9,
                                            frac_of_measured_bf16_sustained=(flop / ms / 1e9) / peaks.get('bf16_tflops_sustained', 1417.2))
    out['rel_l2_vs_fp32_path'] = errs
    print(json.dumps(out))


if __name__ == '__main__':
    main()
