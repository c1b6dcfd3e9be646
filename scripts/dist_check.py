"""torchrun check of the measurement-sharded snapshot gradient (config 5): every rank holds a block
of k-space rows, the partial snapshot gradients are summed with an NCCL all-reduce, the iterates must
match the unsharded run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    # PNP_DIST_BACKEND=gloo: the ranks may share one GPU (NCCL refuses two ranks on a device; gloo stages the CUDA tensor of
    # the all-reduce through the host) -- the multi-rank code path on a one-GPU box (tests/test_gpu_sharded.py)
    backend = os.environ.get('PNP_DIST_BACKEND', 'nccl')
    local = int(os.environ['LOCAL_RANK']) % max(torch.cuda.device_count(), 1)
    torch.cuda.set_device(local)
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend)
    from conftest import rel_l2, synth_image
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    from pnp_svrg_b200.problems import CSMRI
    H = int(os.environ.get('PNP_DIST_SIZE', '512'))
    img = synth_image(H, H, 0)
    kw = dict(eta=0.15 * 0.3 * H * H, T2=5, mini_batch_size=H * H // 40, vr_mode='paper', converge_check=False,
              verbose=False, mb_source='host', mb_seed=3)
    np.random.seed(0)
    p = CSMRI(image=img, H=H, W=H, sample_prob=0.3, snr=20., shard=(rank, world))
    out = pnp_svrg(p, TVDenoiser(), tt=1e9, max_iters=15, **kw)
    np.random.seed(0)
    q = CSMRI(image=img, H=H, W=H, sample_prob=0.3, snr=20.)
    ref = pnp_svrg(q, TVDenoiser(), tt=1e9, max_iters=15, **kw)
    err = rel_l2(out['z'], ref['z'])
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if rank == 0:
        print('sharded snapshot over %d ranks, %dx%d: max rel-L2 vs unsharded = %.3g  PSNR %.2f vs %.2f -> %s'
              % (world, H, H, max(errs), out['psnr_per_iter'][-1], ref['psnr_per_iter'][-1], 'OK' if max(errs) < 1e-5 else 'FAIL'))
    dist.barrier()
    dist.destroy_process_group()
    if max(errs) >= 1e-5:
        sys.exit(1)


if __name__ == '__main__':
    main()
