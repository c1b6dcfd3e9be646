#!/usr/bin/env python
"""bench.py -- PnP-SVRG CSMRI inner iterations/s (BASELINE.json metric) on N GPUs of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on host cores

A "step" is one SVRG epoch on one synthetic image: the snapshot full gradient (mu = grad_full(z),
w = z) followed by T2 inner iterations (minibatch selection, fused gradient + variance-reduced
update, sigma estimate, wavelet prox, PSNR).  value = inner iterations per second, summed over
ranks (each rank reconstructs its own image: independent units, no collective -> weak scaling).

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'pnp_svrg_csmri_inner_iterations_per_s'
UNIT = 'inner_iterations/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--T2', type=int, default=10)
    ap.add_argument('--batch-size', type=int, default=0, help='minibatch size B (default: 100000 at 2048, 1000 at 256)')
    ap.add_argument('--sample-prob', type=float, default=0.3)
    ap.add_argument('--eta', type=float, default=0.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-breakdown', action='store_true')
    ap.add_argument('--cpu-iters', type=int, default=0, help='inner iterations of the CPU sample (default: one epoch)')
    return ap.parse_args()


def workload(a):
    H = a.size
    B = a.batch_size or (100000 if H >= 2048 else max(1, int(1000 * (H / 256.0) ** 2)))
    eta = a.eta or 0.15 * a.sample_prob * H * H          # ~0.15 * M0: stable for the paper-mode step
    return {
        'workload': 'PnP-SVRG (paper-mode VR) CSMRI %dx%d synthetic image, Bernoulli p=%.2f k-space mask, snr 20 dB, '
                    'wavelet-BayesShrink "TV" prox, T2=%d, B=%d, fp32' % (H, H, a.sample_prob, a.T2, B),
        'H': H, 'W': H, 'T2': a.T2, 'mini_batch_size': B, 'sample_prob': a.sample_prob, 'eta': eta,
        'vr_mode': 'paper', 'step': 'one SVRG epoch = snapshot full gradient + T2 inner iterations',
        'l2': 'flushed between steps (256 MiB write, outside the per-step event pairs)',
        'parallelism': 'dp%d independent reconstructions, no collective' % a.gpus,
    }


def make_image(H, seed):
    from conftest import synth_image
    return synth_image(H, H, seed)


# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark(self):
        """timed region starts now: samples taken before this instant are dropped"""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        t_stop = time.time()
        time.sleep(0.05)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        lines = out.strip().splitlines()
        # nvidia-smi -lms samples at a fixed period: keep the tail that falls inside [mark, stop]
        period = 0.02
        keep = max(1, int((t_stop - getattr(self, 't_mark', t_stop)) / period) + 1)
        out = '\n'.join(lines[-keep:])
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v == 'Active':
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------------------
def run_reference(a, cfg, rank, world):
    """The reference's CPU implementation of the path (the NumPy oracle port of its problems/ and
    algorithms/ -- /root/reference itself does not exist on the GPU box), one thread like the
    reference (numpy.fft is single threaded), on a bounded sample of the same workload."""
    if rank != 0:
        return None
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    H, T2, B = cfg['H'], cfg['T2'], cfg['mini_batch_size']
    np.random.seed(0)
    prob = CSMRIPort(make_image(H, 0), H=H, W=H, sample_prob=cfg['sample_prob'], snr=20.)
    per_step = a.cpu_iters or T2
    times = []
    for s in range(a.warmup + a.steps):
        np.random.seed(100 + s)
        t0 = time.time()
        out = AP.pnp_svrg(prob, AP.TVPort(), eta=cfg['eta'], budget=per_step, T2=T2, mini_batch_size=B,
                          vr_mode='paper', converge_check=False)
        dt = time.time() - t0
        if s >= a.warmup:
            times.append(dt)
        prob.Xinit = out['z']
    total = sum(times)
    value = per_step * len(times) / total
    sample = '%d step(s) of %d inner iteration(s) + 1 snapshot gradient each, same image/mask/B' % (len(times), per_step)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': sample,
                         'host_cpus': os.cpu_count(), 'affinity': len(os.sched_getaffinity(0))},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    return line


def cpu_baseline(cfg, iters):
    from oracle import algorithms_port as AP
    from oracle.problems_port import CSMRIPort
    H, T2, B = cfg['H'], cfg['T2'], cfg['mini_batch_size']
    np.random.seed(0)
    prob = CSMRIPort(make_image(H, 0), H=H, W=H, sample_prob=cfg['sample_prob'], snr=20.)
    np.random.seed(1)
    t0 = time.time()
    AP.pnp_svrg(prob, AP.TVPort(), eta=cfg['eta'], budget=iters, T2=T2, mini_batch_size=B, vr_mode='paper',
                converge_check=False)
    dt = time.time() - t0
    return {'value': iters / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': '%d inner iterations + %d snapshot gradient(s) of the same workload, NumPy float64 oracle port, '
                      'numpy.fft single thread as in the reference' % (iters, -(-iters // T2)),
            'seconds': dt, 'host_cpus': os.cpu_count(), 'affinity': len(os.sched_getaffinity(0))}


# --------------------------------------------------------------------------------------------
class Epoch:
    """Device-resident SVRG epoch on the engine primitives (what algorithms.pnp_svrg(fast=True) runs)."""

    def __init__(self, cfg, seed):
        import torch
        from pnp_svrg_b200 import device as D
        from pnp_svrg_b200.denoisers import TVDenoiser
        from pnp_svrg_b200.engine import Engine
        from pnp_svrg_b200.problems import CSMRI
        H = cfg['H']
        np.random.seed(seed)
        self.prob = CSMRI(image=make_image(H, seed), H=H, W=H, sample_prob=cfg['sample_prob'], snr=20.)
        self.den = TVDenoiser()
        self.cfg, self.torch, self.D = cfg, torch, D
        self.eng = Engine(self.prob, self.den, cfg['mini_batch_size'], mb_source='device', mb_seed=seed, fast=True)
        eng = self.eng
        with torch.cuda.stream(eng.stream):
            self.w = torch.empty_like(eng.z)
            self.mu = torch.empty_like(eng.z)
            eng.set_step(cfg['eta'])
        self.n_launch_inner = 0
        self.graph = None

    # the launches of one inner iteration (7 kernels + 1 memset)
    def inner_ops(self, hook=None):
        eng, p, B = self.eng, self.prob, self.cfg['mini_batch_size']
        h = hook or (lambda name: None)
        gk = dict(b=self.w, sel=eng.sel, with_y=False, gscale=1.0 / B, vadd=self.mu, step_ptr=eng.step,
                  z_in=eng.z, z_out=eng.z, clear_sel=True)
        tail = lambda: FUSED_TAIL and p._dev_update_prox(
            1.0 / B, eng.step, self.mu, eng.z, eng.z, eng.sig_log, self.den.sigma_modifier,
            self.den.denoise_strength * self.den.decay ** (self.den.t + 1), p._xrec_dev, eng.mse_log, eng.slot_ptr,
            advance=eng.counters, n_advance=3)
        if hook is None:
            # the minibatch selection only feeds the column pass: draw it on a parallel graph branch
            eng.fork(eng.sample_sel_device, lambda: p._dev_grad(eng.z, phases=1, **gk))
            p._dev_grad(eng.z, phases=2, **gk)
            fused = tail()
            if not fused:
                p._dev_grad(eng.z, phases=4, **gk)
        else:                                  # same kernels, launched one by one so each can be timed
            eng.sample_sel_device(); h('sel_sample')
            p._dev_grad(eng.z, phases=1, **gk); h('lines_r2c')
            p._dev_grad(eng.z, phases=2, **gk); h('cols_mask')
            fused = tail()
            if fused:
                h('c2r+update+prox_fused(sigma+haar+psnr)')
            else:
                p._dev_grad(eng.z, phases=4, **gk); h('lines_c2r+update')
        # kernels launched by this iteration: sel_sample + r2c + cols + [tail] | [c2r + (prox | sigma + haar) + advance]
        if fused:
            self.den.t += 1
            self.n_launch_inner = 4
        elif FUSED_PROX and self.den._dev_prox_fused(self._ctx()):
            h('prox_fused(sigma+haar+psnr)')
            self.n_launch_inner = 6
        else:
            eng.check(eng.lib.pnp_estimate_sigma(self.D.ptr(eng.z), eng.H, eng.W, 1, self.D.ptr(eng.sig_log),
                                                 self.D.ptr(eng.slot_ptr), eng.sptr)); h('sigma_mad')
            self.den._dev_denoise(self._ctx()); h('haar_bayes+psnr')
            self.n_launch_inner = 7
        if not fused:                                      # the fused tail bumps the iteration counters itself
            eng.advance(); h('advance')

    def _ctx(self):
        from pnp_svrg_b200.engine import ProxCtx
        eng = self.eng
        return ProxCtx(eng.z, eng.z, eng.H, eng.W, sig_log=eng.sig_log, xrec=self.prob._xrec_dev,
                       mse_log=eng.mse_log, slot=eng.slot_ptr)

    def snapshot_ops(self):
        eng, p = self.eng, self.prob
        p._dev_grad(eng.z, gscale=1.0 / p.M0, g_out=self.mu)
        eng.copy(self.w, eng.z)

    def capture(self):
        self.graph = self.eng.capture(self.inner_ops)

    def step(self):
        eng = self.eng
        with self.torch.cuda.stream(eng.stream):
            self.snapshot_ops()
            for _ in range(self.cfg['T2']):
                eng.replay(self.graph)
        eng.slot_host += self.cfg['T2']
        if eng.slot_host + self.cfg['T2'] > 4096:
            eng.flush_fast()

    def reset(self):
        eng = self.eng
        with self.torch.cuda.stream(eng.stream):
            eng.z.copy_(self.D.to_lines(self.prob.Xinit, eng.H, eng.W, eng.dev))
            eng._reset_logs()
        eng.stream.synchronize()


FUSED_TAIL = os.environ.get('PNP_BENCH_FUSED_TAIL', '1') == '1'      # pass 3 + update + sigma + wavelet + PSNR as one cooperative launch
FUSED_PROX = os.environ.get('PNP_BENCH_FUSED_PROX', '1') == '1'    # sigma + wavelet + PSNR as one cooperative launch
LAUNCHES_PER_SNAPSHOT = 4     # r2c + cols + c2r + D2D copy


def run_b200(a, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ep = Epoch(cfg, seed=rank)
    eng = ep.eng
    T2, N = cfg['T2'], cfg['H'] * cfg['W']
    with torch.cuda.stream(eng.stream):
        ep.snapshot_ops()
        ep.inner_ops()                     # eager warm iteration before the capture
    eng.stream.synchronize()
    ep.capture()
    ep.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)                    # nvidia-smi needs a moment before its first sample
    for _ in range(a.warmup):
        ep.step()
    barrier()
    if rank == 0:
        sampler.mark()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    t_wall = time.time()
    for s in range(a.steps):
        with torch.cuda.stream(eng.stream):
            flush.fill_(s & 0xff)          # L2 flush, outside the event pair
            evs[s][0].record(eng.stream)
        ep.step()
        evs[s][1].record(eng.stream)
    barrier()
    t_wall = time.time() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = float(sum(ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    psnr = eng.flush_fast()
    value = world * a.steps * T2 / (total_ms * 1e-3)

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': total_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
        'gpu_launches': a.steps * (T2 * ep.n_launch_inner + LAUNCHES_PER_SNAPSHOT),
        'clocks': clocks, 'wall_s_timed_region': t_wall,
        'psnr_first_last': [float(psnr[0]), float(psnr[-1])] if psnr else None,
        'us_per_inner_iteration': 1e3 * total_ms / (a.steps * T2),
    }

    # ---- per-kernel breakdown + roofline (eager launches, CUDA events on the launching stream) ----
    if rank == 0 and not a.no_breakdown:
        line.update(breakdown(a, cfg, ep, 1e3 * total_ms / (a.steps * T2)))

    # ---- end to end through the public API with host buffers ----
    if not a.no_e2e:
        e2e = run_e2e(a, cfg, ep.prob, dev, world, fast=True)
        if e2e is not None:
            eager = run_e2e(a, cfg, ep.prob, dev, world, fast=False)
            e2e['eager_value'] = eager['value']          # same call with fast=False: read-back after every iteration
            line['e2e'] = e2e
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(cfg, a.cpu_iters or T2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


def breakdown(a, cfg, ep, us_inner_graph):
    import torch
    eng = ep.eng
    T2, N = cfg['T2'], cfg['H'] * cfg['W']
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    ep.reset()
    names, events = [], []
    n_iter = min(a.steps * T2, 100)

    def hook(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record(eng.stream)
        events.append((name, e))
    with torch.cuda.stream(eng.stream):
        ep.snapshot_ops()
        for _ in range(n_iter):
            hook('start')
            ep.inner_ops(hook)
    eng.slot_host += n_iter
    eng.stream.synchronize()
    acc = {}
    for (n0, e0), (n1, e1) in zip(events[:-1], events[1:]):
        if n1 == 'start':
            continue
        acc.setdefault(n1, []).append(e0.elapsed_time(e1) * 1e3)
    per = {k: float(np.mean(v)) for k, v in acc.items()}
    tot = sum(per.values())
    # algorithmic (compulsory) bytes per launch, DESIGN.md "Kernels": fp32, N pixels
    alg = {
        'lines_r2c': 12.0 * N,                       # read z, w (8N), write packed half spectrum (4N)
        'cols_mask': 8.5 * N,                        # read + write spectrum (8N), selection bytes (N/2)
        'lines_c2r+update': 16.0 * N,                # read spectrum, mu, z (12N), write z (4N)
        'sigma_mad': 4.0 * N,                        # read z
        'prox_fused(sigma+haar+psnr)': 12.0 * N,     # read z, xrec (8N), write z (4N)
        'c2r+update+prox_fused(sigma+haar+psnr)': 20.0 * N,   # read spectrum, mu, z, xrec (16N), write z (4N)
        'haar_bayes+psnr': 12.0 * N,                 # read z, xrec (8N), write z (4N)
    }
    top = max(alg, key=lambda k: per.get(k, 0.0))
    ach = alg[top] / (per[top] * 1e-6) / 1e9
    traffic = ncu_traffic({'lines_r2c': 'k_lines_r2c', 'cols_mask': 'k_cols_mask', 'lines_c2r+update': 'k_lines_c2r',
                           'sigma_mad': 'k_sigma_mad', 'haar_bayes+psnr': 'k_haar_bayes',
                           'prox_fused(sigma+haar+psnr)': 'k_prox_wavelet_fused',
                           'c2r+update+prox_fused(sigma+haar+psnr)': 'k_update_prox'}[top])
    iter_bytes = 28.125 * N
    out = {
        'kernel_us': per, 'kernel_us_sum_eager': tot,
        'roofline': {'bound': 'hbm', 'kernel': top, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                     'traffic': traffic, 'traffic_source': 'profiles/r01_ncu_full_iteration_kernels.csv (ncu --set full, one launch, cold L2)',
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': alg[top], 'us_per_launch': per[top]},
        'roofline_iteration': {'bound': 'hbm', 'algorithmic_bytes': iter_bytes,
                               'achieved': iter_bytes / (us_inner_graph * 1e-6) / 1e9, 'peak': peak, 'unit': 'GB/s',
                               'frac': iter_bytes / (us_inner_graph * 1e-6) / 1e9 / peak,
                               'note': 'whole inner iteration incl. the amortised snapshot gradient, graph replay'},
    }
    eng.flush_fast()
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu capture"""
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_full_iteration_kernels.csv')
    if not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for r in rows[2:]:
        if kernel in r[0]:
            tot = 0.0
            for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                i = hdr.index(name)
                tot += float(r[i]) * mult.get(units[i], 1.0)
            return tot
    return None


def run_e2e(a, cfg, prob, dev, world, fast=True):
    """Public API, host buffers: algorithms.pnp_svrg with host-drawn minibatches copied from pinned
    memory every inner iteration and the PSNR of every iterate read back (fast=True: graph replay, read-back
    in batches of 64 iterations; fast=False: the reference's loop shape, read-back after every iteration)."""
    import torch
    import torch.distributed as dist
    from pnp_svrg_b200.algorithms import pnp_svrg
    from pnp_svrg_b200.denoisers import TVDenoiser
    T2, B = cfg['T2'], cfg['mini_batch_size']
    iters = a.steps * T2
    kw = dict(eta=cfg['eta'], T2=T2, mini_batch_size=B, vr_mode='paper', verbose=False, converge_check=False,
              mb_source='host', mb_seed=11, fast=fast)
    pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=min(iters, 2 * T2), **kw)          # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.time()
    out = pnp_svrg(prob, TVDenoiser(), tt=1e9, max_iters=iters, **kw)
    torch.cuda.synchronize(dev)
    dt = time.time() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {'value': world * iters / dt, 'unit': UNIT, 'h2d_bytes_per_step': T2 * 4 * B + 4 * prob.N // a.steps,
            'd2h_bytes_per_step': T2 * 16 + 4 * prob.N // a.steps,      # logs + the final iterate (float32 on the wire) 'seconds': dt, 'inner_iterations': iters,
            'api': "pnp_svrg_b200.algorithms.pnp_svrg(problem, denoiser, ..., mb_source='host', fast=%s) -- minibatch drawn "
                   "on the host, copied from pinned memory each inner iteration; PSNR + sigma of every iterate read back "
                   "%s; Xinit upload and final z download included"
                   % (fast, 'in batches of 64 iterations (graph replay)' if fast else 'after every iteration (eager loop)'),
            'psnr_last': float(out['psnr_per_iter'][-1])}


def main():
    a = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = workload(a)
    if a.impl == 'reference':
        line = run_reference(a, cfg, rank, world)
    else:
        line = run_b200(a, cfg, rank, world, local_rank)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
