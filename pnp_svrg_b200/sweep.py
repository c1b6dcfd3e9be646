"""Sweep partitioning over the GPUs of one node -- the data-parallel pattern of the reference's
script_diff_sampratio_set12.py:109-146 / script_diff_snr_set12.py:144-148, which fan a grid of
independent reconstructions out with multiprocessing.Pool.map(process_img, SET12_LIST).

Here: one process per GPU (torchrun), the flat job list (image x sampling ratio x SNR [x algorithm x
denoiser]) is dealt round-robin over the ranks, every rank reconstructs its share on its own GPU, the
per-job records are gathered with all_gather_object.  No collective touches the data path.

    torchrun --nproc-per-node 8 -m pnp_svrg_b200.sweep --images data/Set12 --out sweep.csv
"""
import argparse
import csv
import itertools
import os
import time

import numpy as np

# reference grids: script_diff_sampratio_set12.py:28-29, script_diff_snr_set12.py:29
ALPHAS = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
SNRS = [0., 5., 10., 15., 20., 25., 30.]


def make_jobs(images, alphas=ALPHAS, snrs=SNRS, problems=('CSMRI',), denoisers=('TV',), algos=('pnp_svrg',)):
    """Flat, deterministic job list; job['id'] is its position."""
    jobs = []
    for i, (img, prob, den, algo, a, s) in enumerate(itertools.product(images, problems, denoisers, algos, alphas, snrs)):
        jobs.append(dict(id=i, image=img, problem=prob, denoiser=den, algo=algo, alpha=a, snr=s))
    return jobs


def partition(jobs, rank, world):
    """Round-robin deal: rank r takes jobs r, r + world, r + 2*world, ..."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return jobs[rank::world]


def run_partitioned(jobs, runner, rank=0, world=1, gather=True):
    """Run this rank's share with ``runner(job) -> dict``; returns all records sorted by job id on
    every rank (when gather and torch.distributed is initialised), else the local ones."""
    local = []
    for job in partition(jobs, rank, world):
        try:
            rec = runner(job)
        except Exception as e:             # a failed job must not take the sweep down (Pool.map would)
            rec = dict(error=repr(e))
        rec = dict(rec)
        rec.setdefault('id', job['id'])
        rec['rank'] = rank
        local.append(rec)
    if gather and world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('world > 1 needs an initialised torch.distributed process group')
        parts = [None] * world
        dist.all_gather_object(parts, local)
        local = [r for part in parts for r in part]
    return sorted(local, key=lambda r: r['id'])


def reconstruct(job, H=256, W=256, eta_scale=0.15, T2=10, mini_batch_size=1000, iters=200, images=None, seed=0):
    """One reconstruction of the sweep on the current GPU (fixed hyper-parameters, no hyperopt)."""
    from . import algorithms as ALG
    from . import denoisers as DN
    from . import problems as PR
    img = job['image']
    kw = dict(image=images[img]) if images is not None and not isinstance(img, str) else dict(img_path=img)
    np.random.seed(seed + job['id'])
    if job['problem'] != 'CSMRI':
        raise NotImplementedError('sweep runner: only CSMRI jobs are wired up in this revision')
    p = PR.CSMRI(H=H, W=W, sample_prob=job['alpha'], snr=job['snr'], **kw)
    den = {'TV': DN.TVDenoiser, 'NLM': DN.NLMDenoiser}[job['denoiser']]()
    B = min(mini_batch_size, p.M0)
    t0 = time.time()
    eta = min(eta_scale * p.M0, 3.0 * B)          # full-gradient step ~0.15 * M0, capped by the minibatch term
    kw = dict(eta=eta, tt=1e9, verbose=False, converge_check=False, max_iters=iters, fast=True, sync_every=iters)
    if job['algo'] != 'pnp_gd':
        kw.update(mini_batch_size=B, mb_source='device', mb_seed=job['id'])
    if job['algo'] in ('pnp_svrg', 'pnp_sarah'):
        kw['T2'] = T2
    if job['algo'] == 'pnp_svrg':
        kw['vr_mode'] = 'paper'
    out = getattr(ALG, job['algo'])(p, den, **kw)
    dt = time.time() - t0
    ps = out['psnr_per_iter']
    return dict(id=job['id'], image=str(img), alpha=job['alpha'], snr=job['snr'], algo=job['algo'],
                denoiser=job['denoiser'], psnr_init=float(ps[0]), psnr_final=float(ps[-1]), iters=iters, seconds=dt)


_RUNNERS = {}        # reusable batched engines of device-built sweeps, keyed by shape


def reconstruct_batch(jobs, H=256, W=256, eta_scale=0.15, T2=10, mini_batch_size=1000, iters=200, images=None,
                      seed=0, host_threads=8, construct='host'):
    """A group of same-size CSMRI jobs as ONE batched device run (pnp_svrg_b200.batched).
    construct='host': problems built with NumPy in the reference's draw order (thread pool);
    construct='device': the whole batch is built on the GPU (batched.csmri_device_batch) -- the sweeps are
    unseeded in the reference, so only the distribution of the draws matters there."""
    from concurrent.futures import ThreadPoolExecutor
    from .batched import BatchedSVRG, csmri_device_batch, csmri_host_spec
    from .problems.problem import load_image

    def image_of(job):
        img = images[job['image']] if images is not None and not isinstance(job['image'], str) else None
        if img is None:
            from PIL import Image
            img = np.array(Image.open(job['image']).resize((H, W)))
        return img
    if construct == 'device':
        t0 = time.time()
        batch = csmri_device_batch([image_of(j) for j in jobs], [j['alpha'] for j in jobs], [j['snr'] for j in jobs], H, W,
                                   seed=seed + jobs[0]['id'])
        m0 = batch['m0_host']
        B = int(min(mini_batch_size, m0.min()))
        etas = [min(eta_scale * float(m), 3.0 * B) for m in m0]
        key = (len(jobs), H, W, B, T2, iters)
        run = _RUNNERS.get(key)                      # same shape as an earlier batch: reuse buffers and the captured graph
        if run is None:
            run = _RUNNERS[key] = BatchedSVRG(batch, T2=T2, mini_batch_size=B, etas=etas, seed=seed, max_slots=iters)
            run.whole_run_graph = True
            if run.sup_stride == run.N:
                run.prepare_build()                  # later groups are built straight into this engine (build_from_images)
        else:
            run.reload(batch, etas)
        run.run(iters)
        out = run.results(with_z=False)
        dt = time.time() - t0
        return [dict(id=j['id'], image=str(j['image']), alpha=j['alpha'], snr=j['snr'], algo='pnp_svrg', denoiser='TV',
                     psnr_init=float(out['psnr_init'][i]), psnr_final=float(out['psnr'][-1, i]), iters=iters,
                     seconds=dt / len(jobs)) for i, j in enumerate(jobs)]
    if construct != 'host':
        raise ValueError("construct must be 'host' or 'device'")

    def spec(job):
        img = images[job['image']] if images is not None and not isinstance(job['image'], str) else None
        if img is None:
            from PIL import Image
            img = np.array(Image.open(job['image']).resize((H, W)))
        return csmri_host_spec(img, H, W, job['alpha'], job['snr'], rng=np.random.RandomState(seed + job['id']))
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=host_threads) as ex:
        specs = list(ex.map(spec, jobs))
    B = min([mini_batch_size] + [s['M0'] for s in specs])
    etas = [min(eta_scale * s['M0'], 3.0 * B) for s in specs]
    run = BatchedSVRG(specs, T2=T2, mini_batch_size=B, etas=etas, seed=seed + jobs[0]['id'], max_slots=iters)
    run.run(iters)
    out = run.results()
    run.close()
    dt = time.time() - t0
    return [dict(id=j['id'], image=str(j['image']), alpha=j['alpha'], snr=j['snr'], algo='pnp_svrg', denoiser='TV',
                 psnr_init=float(out['psnr_init'][i]), psnr_final=float(out['psnr'][-1, i]), iters=iters,
                 seconds=dt / len(jobs)) for i, j in enumerate(jobs)]


class DeviceBatchPipeline:
    """Device-built sweep batches, double buffered: while the batched engine of group k runs (its own stream), group
    k + 1 is built on the current stream (batched.csmri_device_batch: mask, measurements, Xinit, support lists) and
    loaded into the OTHER engine; the host only waits for a run when it needs its engine back or its records.
    ``submit(group)`` enqueues a group and returns the records of the group that finished meanwhile (or None);
    ``drain()`` returns the rest.  Same arguments and records as ``reconstruct_batch(..., construct='device')``."""

    def __init__(self, H=256, W=256, eta_scale=0.15, T2=10, mini_batch_size=1000, iters=200, images=None, seed=0, depth=2):
        self.kw = dict(H=H, W=W, eta_scale=eta_scale, T2=T2, mini_batch_size=mini_batch_size, iters=iters, images=images, seed=seed)
        self.depth = depth
        self.engines = {}            # (slot, shape key) -> BatchedSVRG
        self.inflight = []           # (slot, run, jobs, t0)
        self.n = 0
        self.batch_size = None       # size of the first group: shorter groups (the tail of a rank's share) are padded to it
        self.build_seconds = 0.0     # host time inside csmri_device_batch (ends with a device->host read of M0: ~ its GPU time)

    def _image(self, job):
        images, H, W = self.kw['images'], self.kw['H'], self.kw['W']
        img = images[job['image']] if images is not None and not isinstance(job['image'], str) else None
        if img is None:
            from PIL import Image
            img = np.array(Image.open(job['image']).resize((H, W)))
        return img

    def _collect(self, entry):
        slot, run, jobs, t0 = entry
        jobs = [j for j in jobs if not j.get('_pad')]
        out = run.results(with_z=False)
        dt = time.time() - t0
        iters = self.kw['iters']
        if (out['m0'][:len(jobs)] < run.B).any():
            # the batches are built without a read-back, so this is the first place M0 is seen on the host
            raise ValueError('mini_batch_size %d exceeds the number of measurements of a problem (min M0 = %d)'
                             % (run.B, int(out['m0'].min())))
        return [dict(id=j['id'], image=str(j['image']), alpha=j['alpha'], snr=j['snr'], algo='pnp_svrg', denoiser='TV',
                     psnr_init=float(out['psnr_init'][i]), psnr_final=float(out['psnr'][-1, i]), iters=iters,
                     seconds=dt / len(jobs)) for i, j in enumerate(jobs)]

    def submit(self, jobs):
        from .batched import BatchedSVRG, csmri_device_batch
        k = self.kw
        t0 = time.time()
        if self.batch_size is None:
            self.batch_size = len(jobs)
        elif len(jobs) < self.batch_size:
            # same shape as the engines that exist (no allocation, no new launch geometry inside a sweep): repeat the last job;
            # the copies' records are dropped
            jobs = list(jobs) + [dict(jobs[-1], _pad=True)] * (self.batch_size - len(jobs))
        done = None
        if len(self.inflight) >= self.depth:         # the engine this group needs: its previous run finished `depth` groups ago
            done = self._collect(self.inflight.pop(0))
        slot = self.n % self.depth
        B = int(k['mini_batch_size'])
        run = self.engines.get((slot, len(jobs), B))
        if run is not None and os.environ.get('PNP_BUILD_TORCH', '0') != '1' and not any(e[1] is run for e in self.inflight):
            # the engine exists and is idle: the package's own constructor writes the group straight into its buffers on its
            # stream (no allocation, no copy, no read-back); it overlaps the run of the other engine
            tb = time.time()
            run.build_from_images([self._image(j) for j in jobs], [j['alpha'] for j in jobs], [j['snr'] for j in jobs],
                                  k['seed'] + jobs[0]['id'], k['eta_scale'], 3.0 * B)
            self.build_seconds += time.time() - tb
            run.run(k['iters'])
            self.inflight.append((slot, run, jobs, t0))
            self.n += 1
            return done
        tb = time.time()
        batch = csmri_device_batch([self._image(j) for j in jobs], [j['alpha'] for j in jobs], [j['snr'] for j in jobs], k['H'], k['W'],
                                   seed=k['seed'] + jobs[0]['id'], sync=False)
        self.build_seconds += time.time() - tb
        m0 = batch['m0_host']
        if m0 is None:
            # built without a read-back (the package's own constructor): M0 stays on the device, the step sizes are derived
            # from it there, and B <= M0 is checked when the records come back (_collect)
            import torch
            B = int(k['mini_batch_size'])
            etas = torch.clamp(batch['m0'].to(torch.float32) * float(k['eta_scale']), max=3.0 * B)
        else:
            B = int(min(k['mini_batch_size'], m0.min()))
            etas = [min(k['eta_scale'] * float(m), 3.0 * B) for m in m0]
        key = (slot, len(jobs), B)
        run = self.engines.get(key)
        if run is None:
            run = self.engines[key] = BatchedSVRG(batch, T2=k['T2'], mini_batch_size=B, etas=etas, seed=k['seed'], max_slots=k['iters'])
            run.whole_run_graph = True
        else:
            if any(e[1] is run for e in self.inflight):                  # same engine still in flight (depth 1 or odd shapes)
                i = next(i for i, e in enumerate(self.inflight) if e[1] is run)
                extra = self._collect(self.inflight.pop(i))
                done = (done or []) + extra
            run.reload(batch, etas)
        run.run(k['iters'])
        self.inflight.append((slot, run, jobs, t0))
        self.n += 1
        return done

    def drain(self):
        out = []
        while self.inflight:
            out.extend(self._collect(self.inflight.pop(0)))
        return out

    def close(self):
        for run in self.engines.values():
            run.close()
        self.engines = {}


def run_partitioned_batched(jobs, batch_runner, rank=0, world=1, batch=56, gather=True):
    """Like run_partitioned, but this rank's share is processed in groups of `batch` jobs.  ``batch_runner`` is a callable
    group -> records, or a pipeline object with submit(group) / drain() (DeviceBatchPipeline)."""
    mine = partition(jobs, rank, world)
    local = []
    if hasattr(batch_runner, 'submit'):
        t0 = time.time()
        for k in range(0, len(mine), batch):
            recs = batch_runner.submit(mine[k:k + batch])
            if recs:
                local.extend(recs)
        local.extend(batch_runner.drain())
        t1 = time.time()
        for r in local:
            r['rank'] = rank
        if gather and world > 1:
            local = _gather_records(jobs, local, rank, world)
        # where this rank's wall clock went: its own groups (build + run + read-back) / the record gather (incl. waiting for
        # the slowest rank)
        batch_runner.timing = dict(groups_seconds=t1 - t0, gather_seconds=time.time() - t1)
        return sorted(local, key=lambda r: r['id'])
    for k in range(0, len(mine), batch):
        group = mine[k:k + batch]
        try:
            recs = batch_runner(group)
        except Exception as e:
            recs = [dict(id=j['id'], error=repr(e)) for j in group]
        for r in recs:
            r['rank'] = rank
        local.extend(recs)
    if gather and world > 1:
        local = _gather_records(jobs, local, rank, world)
    return sorted(local, key=lambda r: r['id'])


def _gather_records(jobs, local, rank, world):
    """All ranks' records on every rank.  The numeric part (id, PSNRs, seconds, rank) travels as one padded tensor
    all-gather; the descriptive fields are rebuilt from the job list every rank holds.  Records that carry an error
    message fall back to the pickling collective (rare, and strings do not fit a tensor)."""
    import torch
    import torch.distributed as dist
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    cap = (len(jobs) + world - 1) // world
    # one collective: row `cap` carries this rank's "some record holds an error message" flag, so the common case needs no
    # separate all-reduce (and no device -> host read) before the gather
    bad = any('error' in r for r in local)
    rows = np.full((cap + 1, 5), -1.0)
    if local and not bad:
        rows[:len(local)] = [(r['id'], r['psnr_init'], r['psnr_final'], r['seconds'], r['rank']) for r in local]
    rows[cap, 0] = 1.0 if bad else 0.0
    t = torch.from_numpy(rows).to(dev)
    allrows = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(allrows.view(-1, 5), t)
    allrows = allrows.cpu().numpy()
    if (allrows[:, cap, 0] > 0).any():
        parts = [None] * world
        dist.all_gather_object(parts, local)
        return [r for part in parts for r in part]
    by_id = {j['id']: j for j in jobs}
    proto = local[0] if local else {}
    algo, den, iters = proto.get('algo'), proto.get('denoiser'), proto.get('iters')
    rows = allrows[:, :cap].reshape(-1, 5)
    rows = rows[rows[:, 0] >= 0]
    out = []
    # (plain Python values from two tolist() calls: 840 records are ~1 ms of the host's time on the sweep's critical path)
    for jid, (p0, p1, sec, rk) in zip(rows[:, 0].astype(np.int64).tolist(), rows[:, 1:].tolist()):
        j = by_id[jid]
        out.append({'id': jid, 'image': str(j['image']), 'alpha': j['alpha'], 'snr': j['snr'], 'algo': algo or j.get('algo'),
                    'denoiser': den or j.get('denoiser'), 'psnr_init': p0, 'psnr_final': p1, 'iters': iters, 'seconds': sec,
                    'rank': int(rk)})
    return out


# ------------------------------------------------------------------------------------------ tuning sweeps
# The reference's sweep scripts do not run fixed hyper-parameters: every grid cell is a hyperopt search over the
# tune_pnp_* objective (script_diff_sampratio_set12.py:41-134, script_diff_snr_set12.py likewise).  The helpers
# below keep that shape -- get_problem / get_denoiser / get_proxy_pspace / one CSV row per cell -- on top of
# pnp_svrg_b200.search, so a cell is `tune_job(job)` and a sweep is run_partitioned(jobs, tune_job, rank, world).
TIME_PER_TRIAL = 30               # script_diff_sampratio_set12.py:34-35
MAX_EVALS = 100
ETA_RANGE, MB_RANGE, T2_RANGE, DSTR_RANGE = (0, 100), (1, 100), (1, 100), (0, 2)      # :37-40
TUNE_ARGS = {                     # order of the objective's positional tuple (algorithms/pnp_*.py tune_pnp_*)
    'pnp_gd': ('eta', 'dstrength'),
    'pnp_sgd': ('eta', 'mini_batch_size', 'dstrength'),
    # the reference's sweep script builds a 3-tuple here while tune_pnp_saga unpacks four values (every SAGA trial of the
    # original raises ValueError); the table size is searched like the other integer parameters
    'pnp_saga': ('eta', 'mini_batch_size', 'dstrength', 'hist_size'),
    'pnp_sarah': ('eta', 'mini_batch_size', 'T2', 'dstrength'),
    'pnp_svrg': ('eta', 'mini_batch_size', 'T2', 'dstrength'),
}


def get_problem(prob_name, im_path, alpha, snr, H=256, W=256, kernel='Minimal', image=None, pr_size=32):
    """script_diff_sampratio_set12.py:41-51.  ``alpha`` is the sampling ratio as a fraction in (0, 1] (the
    reference's ALPHA_LIST holds 1..10 and divides by ten); PR runs at 32 x 32 there (dense A)."""
    from . import problems as PR
    kw = dict(image=image) if image is not None else dict(img_path=im_path)
    if prob_name == 'CSMRI':
        return PR.CSMRI(H=H, W=W, sample_prob=alpha, snr=snr, **kw)
    if prob_name == 'DeblurSR':
        return PR.Deblur(kernel=kernel, H=H, W=W, scale_percent=int(round(alpha * 100)), snr=snr, **kw)
    if prob_name == 'PR':
        return PR.PhaseRetrieval(H=pr_size, W=pr_size, num_meas=int(alpha * 10 * pr_size * pr_size), snr=snr, **kw)
    raise Exception('Problem name "{0}" not found'.format(prob_name))


def get_denoiser(dnr_name):
    """script_diff_sampratio_set12.py:53-63 (its 'CNN' branch names a class that does not exist there)."""
    from . import denoisers as DN
    table = {'BM3D': DN.BM3DDenoiser, 'NLM': DN.NLMDenoiser, 'TV': DN.TVDenoiser}
    if dnr_name not in table:
        raise Exception('Denoiser name "{0}" not found'.format(dnr_name))
    return table[dnr_name]()


def get_pspace(algo_name, eta=ETA_RANGE, mb=MB_RANGE, T2=T2_RANGE, dstr=DSTR_RANGE):
    """The search space of one algorithm as the tuple its tune_pnp_* objective unpacks (:65-107)."""
    from .search import hp, quniform, scope
    if algo_name not in TUNE_ARGS:
        raise Exception('Algorithm name "{0}" not found'.format(algo_name))
    nodes = {'eta': lambda: hp.uniform('eta', *eta),
             'mini_batch_size': lambda: scope.int(quniform('mini_batch_size', mb[0], mb[1], q=1)),
             'T2': lambda: scope.int(quniform('T2', T2[0], T2[1], q=1)),
             'dstrength': lambda: hp.uniform('dstrength', *dstr),
             'hist_size': lambda: scope.int(quniform('hist_size', 1, 100, q=1))}
    return tuple(nodes[k]() for k in TUNE_ARGS[algo_name])


def get_proxy_pspace(main_problem, algo_name, denoiser, tt=TIME_PER_TRIAL, spaces=None, **extra):
    """(objective, space) of one grid cell; ``extra`` goes to the loop (max_iters=, fast=, mb_source=, ...)."""
    from functools import partial
    from . import algorithms as ALG
    pspace = get_pspace(algo_name, **(spaces or {}))
    proxy = partial(getattr(ALG, 'tune_' + algo_name), problem=main_problem, denoiser=denoiser, tt=tt, verbose=False,
                    lr_decay=1, converge_check=True, diverge_check=True, **extra)
    return proxy, pspace


def tuning_row(job, trials, best):
    """One CSV row in the reference's layout (:129-133): problem, denoiser, algorithm, alpha, snr, best loss,
    'PARAMETERS:', then label, value pairs."""
    row = [job['problem'], job['denoiser'], job['algo'], job['alpha'], job['snr'],
           trials.best_trial['result']['loss'], 'PARAMETERS:']
    for key in best:
        row.append(key)
        row.append(best[key])
    return row


def tune_job(job, max_evals=MAX_EVALS, tt=TIME_PER_TRIAL, seed=0, H=256, W=256, images=None, spaces=None, algo='tpe',
             **extra):
    """Hyper-parameter search of one grid cell on the current GPU; returns a record whose 'row' is the CSV row."""
    from . import search
    img = job['image']
    image = images[img] if images is not None and not isinstance(img, str) else None
    np.random.seed(seed + job['id'])
    p = get_problem(job['problem'], img if image is None else None, job['alpha'], job['snr'], H=H, W=W, image=image)
    dnr = get_denoiser(job['denoiser'])
    space_kw = dict(spaces or {})
    if 'mb' not in space_kw:                         # a minibatch cannot exceed the measurements there are
        cap = int(getattr(p, 'M0', p.M))
        space_kw['mb'] = (MB_RANGE[0], max(MB_RANGE[0] + 1, min(MB_RANGE[1], cap)))
    proxy, pspace = get_proxy_pspace(p, job['algo'], dnr, tt=tt, spaces=space_kw, **extra)
    trials = search.Trials()
    t0 = time.time()
    best = search.fmin(proxy, space=pspace, algo={'tpe': search.tpe.suggest, 'rand': search.rand.suggest}[algo],
                       trials=trials, max_evals=max_evals, rstate=np.random.default_rng(seed + job['id']), catch=True)
    return dict(id=job['id'], image=str(img), problem=job['problem'], denoiser=job['denoiser'], algo=job['algo'],
                alpha=job['alpha'], snr=job['snr'], loss=float(trials.best_trial['result']['loss']), best=dict(best),
                trials=len(trials), seconds=time.time() - t0, row=tuning_row(job, trials, best))


def write_tuning_csv(path, records):
    """The sweep scripts' output file (:155-160): a 'Results:' line, then one row per grid cell."""
    with open(path, 'w', newline='') as f:
        w = csv.writer(f, delimiter=',')
        w.writerow(['Results:'])
        for r in records:
            if 'row' in r:
                w.writerow(r['row'])


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', required=True, help='directory of grey images (e.g. data/Set12)')
    ap.add_argument('--out', default='sweep.csv')
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--tune', type=int, default=0, metavar='MAX_EVALS',
                    help='search the hyper-parameters of every grid cell with this many trials (the reference scripts use 100) '
                         'and write the tuning CSV instead of fixed-parameter reconstructions')
    ap.add_argument('--tt', type=float, default=TIME_PER_TRIAL, help='wall-clock budget of one trial in seconds')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    files = sorted(os.path.join(a.images, f) for f in os.listdir(a.images) if f.lower().endswith(('.png', '.jpg')))
    jobs = make_jobs(files)
    t0 = time.time()
    if a.tune > 0:
        recs = run_partitioned(jobs, lambda j: tune_job(j, max_evals=a.tune, tt=a.tt, H=a.size, W=a.size), rank, world)
        if rank == 0:
            write_tuning_csv(a.out, recs)
            print('%d grid cells tuned (%d trials each) in %.1f s on %d GPU(s)' % (len(recs), a.tune, time.time() - t0, world))
        if world > 1:
            dist.destroy_process_group()
        return
    recs = run_partitioned(jobs, lambda j: reconstruct(j, H=a.size, W=a.size, iters=a.iters), rank, world)
    if rank == 0:
        with open(a.out, 'w', newline='') as f:
            wr = csv.DictWriter(f, fieldnames=sorted({k for r in recs for k in r}))
            wr.writeheader()
            wr.writerows(recs)
        print('%d reconstructions in %.1f s on %d GPU(s): %.2f recon/s' % (len(recs), time.time() - t0, world,
                                                                          len(recs) / (time.time() - t0)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
