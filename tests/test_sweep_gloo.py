"""CPU, world_size 2 over gloo: the sweep's partition / gather logic (the N>1 path of config 4)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from pnp_svrg_b200 import sweep
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    jobs = sweep.make_jobs(['a.png', 'b.png', 'c.png'], alphas=[0.1, 0.5], snrs=[0., 20., 30.])

    def runner(job):
        if job['id'] == 4:
            raise RuntimeError('boom')                 # a failing job is recorded, not fatal
        return dict(id=job['id'], value=job['id'] ** 2, alpha=job['alpha'])
    recs = sweep.run_partitioned(jobs, runner, rank, world)
    q.put((rank, recs))
    dist.barrier()
    dist.destroy_process_group()


def _worker_batched(rank, world, port, q, fail):
    sys.path.insert(0, ROOT)
    from pnp_svrg_b200 import sweep
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    jobs = sweep.make_jobs(['a.png', 'b.png', 'c.png'], alphas=[0.1, 0.5], snrs=[0., 20., 30.])[:17]     # ragged shares

    def batch_runner(group):
        if fail and any(j['id'] == 6 for j in group):
            raise RuntimeError('boom')
        return [dict(id=j['id'], image=str(j['image']), alpha=j['alpha'], snr=j['snr'], algo='pnp_svrg', denoiser='TV',
                     psnr_init=10.0 + j['id'], psnr_final=20.0 + j['id'], iters=7, seconds=0.5) for j in group]
    recs = sweep.run_partitioned_batched(jobs, batch_runner, rank, world, batch=4)
    q.put((rank, recs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('fail', [False, True])
def test_world_size_2_gloo_batched_gather(fail):
    """run_partitioned_batched: numeric records travel as one tensor all-gather (pickling fallback on errors)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + int(fail)
    procs = [ctx.Process(target=_worker_batched, args=(r, 2, port, q, fail)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]
    recs = got[0]
    assert [r['id'] for r in recs] == list(range(17))
    assert all(r['rank'] == r['id'] % 2 for r in recs)
    ok = [r for r in recs if 'error' not in r]
    assert all(r['psnr_final'] == 20.0 + r['id'] and r['psnr_init'] == 10.0 + r['id'] and r['iters'] == 7 and r['denoiser'] == 'TV'
               for r in ok)
    assert len(ok) == (17 if not fail else 13) and (not fail or all('boom' in r['error'] for r in recs if 'error' in r))


def test_partition_covers_every_job_once():
    from pnp_svrg_b200 import sweep
    jobs = sweep.make_jobs(['x'] * 12)
    assert len(jobs) == 12 * 10 * 7 == 840              # config 4: 12 images x 10 ratios x 7 SNRs
    for world in (1, 2, 4, 8):
        seen = sorted(j['id'] for r in range(world) for j in sweep.partition(jobs, r, world))
        assert seen == list(range(840))
        sizes = [len(sweep.partition(jobs, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sweep.partition(jobs, 2, 2)


def test_world_size_2_gloo_gather():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                               # every rank ends with the full, ordered table
    recs = got[0]
    assert [r['id'] for r in recs] == list(range(18))
    assert all(r['rank'] == r['id'] % 2 for r in recs)
    assert 'error' in recs[4] and 'boom' in recs[4]['error']
    assert all(r['value'] == r['id'] ** 2 for r in recs if r['id'] != 4)


def test_measurement_shards_partition_the_rows():
    """config 5: the packed half-spectrum rows are split into disjoint contiguous bands that cover [0, H/2), and every
    k-space row (with its Hermitian mirror, and the Nyquist row with DC) falls into exactly one band."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location('csmri_mod', os.path.join(ROOT, 'pnp_svrg_b200', 'problems', 'CSMRI.py'))
    src = open(spec.origin).read()
    ns = {'np': np}
    exec(src[src.index('def shard_rows'):src.index('class CSMRI')], ns)
    for H in (32, 256, 2048):
        for world in (1, 2, 4, 8, 3):
            blocks = [ns['shard_rows'](H, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == H // 2
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            kyp = ns['packed_row_of'](np.arange(H), H)
            assert kyp.min() == 0 and kyp.max() == H // 2 - 1
            assert np.array_equal(kyp[1:H // 2], kyp[H - 1:H // 2:-1])          # ky and H - ky share a packed row
            owner = np.full(H, -1)
            for r, (lo, hi) in enumerate(blocks):
                sel = (kyp >= lo) & (kyp < hi)
                assert (owner[sel] == -1).all()
                owner[sel] = r
            assert (owner >= 0).all()
