"""GPU: measurement-sharded snapshot gradient (config 5).  Single process: the sum of the per-shard
partial gradients equals the full gradient; with 2 visible GPUs the NCCL path is exercised too."""
import numpy as np
import pytest

from conftest import rel_l2, synth_image

pytestmark = pytest.mark.gpu


def test_partial_gradients_sum_to_full(cuda):
    from pnp_svrg_b200.problems import CSMRI
    img = synth_image(128, 128, 1)
    z = np.random.default_rng(0).uniform(0, 1, 128 * 128)
    np.random.seed(0)
    full = CSMRI(image=img, H=128, W=128, sample_prob=0.3, snr=20.)
    g = full.grad_full(z)
    for world in (2, 4):
        acc = np.zeros_like(g)
        cnt = 0
        for r in range(world):
            np.random.seed(0)
            p = CSMRI(image=img, H=128, W=128, sample_prob=0.3, snr=20., shard=(r, world))
            assert np.array_equal(p.mask, full.mask) and p.M0 == full.M0        # replicas of the same problem
            acc += p.grad_full(z)                                                # partial (divided by the GLOBAL M0)
            cnt += p._shard_count
        assert cnt == full.M0
        assert rel_l2(acc, g) < 2e-6


def test_pr_row_block_partial_gradients_sum_to_full(cuda):
    """problems/PR.py:75-79 with the rows of A cut into one block per rank (PhaseRetrieval(shard=)): the partial
    gradients (each already divided by the global M) add up to the unsharded full gradient; the snapshot of pnp_svrg
    goes through the same partial sums."""
    from pnp_svrg_b200.problems import PhaseRetrieval
    from pnp_svrg_b200.problems.PR import shard_block
    img = synth_image(32, 32, 2)
    z = np.random.default_rng(1).uniform(0.1, 1, 32 * 32)
    np.random.seed(0)
    full = PhaseRetrieval(image=img, H=32, W=32, num_meas=1500, snr=20.)
    g = full.grad_full(z)
    for world in (2, 3, 8):
        acc = np.zeros_like(g)
        rows = 0
        for r in range(world):
            np.random.seed(0)
            p = PhaseRetrieval(image=img, H=32, W=32, num_meas=1500, snr=20., shard=(r, world))
            assert np.array_equal(p.A, full.A)                                 # replicas of the same problem
            lo, hi = shard_block(1500, r, world)
            assert p._shard_block == (lo, hi)
            rows += hi - lo
            acc += p.grad_full(z)
            with pytest.raises(NotImplementedError):
                p._dev_grad(p._r, g_out=p._r)                                  # a silent partial sum is refused
        assert rows == 1500
        assert rel_l2(acc, g) < 2e-6, (world, rel_l2(acc, g))


def test_pr_split_column_pass_equals_the_unsplit_one(cuda):
    """pnp_pr_grad with and without the row-chunk scratch (the transposed product split over rows as well as columns):
    same gradient to rounding of the summation order, full and minibatch forms, two-point form included."""
    import ctypes as C
    import torch
    from pnp_svrg_b200 import _lib, device as D
    from pnp_svrg_b200.problems import PhaseRetrieval
    np.random.seed(3)
    p = PhaseRetrieval(image=synth_image(32, 32, 5), H=32, W=32, num_meas=777, snr=25.)
    dev = p._device
    rng = np.random.default_rng(0)
    z = D.to_lines(rng.uniform(0.1, 1, p.N), 32, 32, dev)
    w = D.to_lines(rng.uniform(0.1, 1, p.N), 32, 32, dev)
    for sel in (None, torch.from_numpy(rng.choice(777, 50, replace=False).astype(np.int32)).to(dev)):
        outs = []
        for chunks in (0, 1, 7, 64):
            g = torch.zeros_like(z)
            args = _lib.PrGradArgs(A=D.ptr(p._A), n=p.N, M=int(p.M), z=D.ptr(z), w=D.ptr(w), y=D.ptr(p._y), rows=D.ptr(sel),
                                   count=0 if sel is None else int(sel.numel()), cursor=None, r=D.ptr(p._r), gscale=0.01,
                                   step=0.0, step_ptr=None, g_out=D.ptr(g), vadd=None, v_out=None, z_in=None, z_out=None,
                                   partial=D.ptr(p._partial) if chunks else None, partial_chunks=chunks)
            _lib.check(_lib.load().pnp_pr_grad(C.byref(args), D.stream()))
            outs.append(g.cpu().numpy())
        for o in outs[1:]:
            assert rel_l2(o, outs[0]) < 1e-6


def test_two_ranks_sharded_snapshot_allreduce(cuda):
    """config 5 with TWO PROCESSES: every rank holds the measurements of its band of k-space rows, the snapshot gradients of
    pnp_svrg are summed by torch.distributed.all_reduce (CSMRI._snapshot_allreduce), the iterates match the unsharded run
    (scripts/dist_check.py, launched through torch.distributed.run).  NCCL when two GPUs are visible; on a one-GPU box both
    ranks use cuda:0 and the collective runs over gloo (NCCL refuses two ranks on one device) -- same code path above it."""
    import os
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    env = dict(os.environ, PNP_DIST_BACKEND=backend, PNP_DIST_SIZE='256')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(root, 'scripts', 'dist_check.py')]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert 'OK' in r.stdout and 'sharded snapshot over 2 ranks' in r.stdout, r.stdout[-1000:]
