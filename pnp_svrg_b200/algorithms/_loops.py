"""The five PnP loops on the device engine.  Control flow, logging and stop rules follow the
reference line by line (cited per function); the arithmetic is the fused CUDA path.

Additive keyword arguments (all optional, defaults reproduce the reference behaviour):
  max_iters   stop after this many prox (denoiser) calls, in addition to the wall-clock ``tt``
  vr_mode     pnp_svrg only: 'as_committed' (v = mu, algorithms/pnp_svrg.py:54) or
              'paper' (v = (g_B(z) - g_B(w))/B + mu, the commented line :53)
  mb_source   'legacy' np.random global RNG in the reference's call order (default),
              'host' np.random.Generator(mb_seed), 'device' GPU sampler, 'stream' use mb_stream
  fast        replay a captured CUDA graph per inner iteration and defer the PSNR read-back to the
              end / every ``sync_every`` iterations (stop rules are then applied at those points)
"""
import os
import time

import numpy as np
import torch

from .. import device as D
from ..engine import LOG_CHUNK, Budget, Engine, stop_rule


def _grad_update(eng, a, b, sel, with_y, gscale, **kw):
    # the engine's minibatch selection is single use: the pass that consumes it leaves it zeroed
    if eng.chain:
        kw['chain'] = True
    eng.p._dev_grad(a, b=b, sel=sel, with_y=with_y, gscale=gscale, clear_sel=sel is not None and sel is eng.sel, **kw)


def _grad_update_prox(eng, a, b, sel, gscale, vadd, z, advance=0, sel_fn=None, zero_grad=False, r2c_done=False, next_pass=None):
    """z <- prox(z - step * (g_sel(a - b) * gscale + vadd)).  CSMRI + wavelet prox: the inverse line pass, the update,
    the sigma estimate and the prox run as ONE cooperative launch on lines resident in shared memory
    (pnp_csmri_update_prox); otherwise the gradient pass followed by ``eng.prox``.  ``advance`` > 0 also bumps
    that many end-of-iteration counters (``eng.advance``), inside the same launch when it is the fused one.
    ``sel_fn`` rebuilds the minibatch selection; on the CSMRI path the forward line pass does it itself
    (``Engine.sel_job``: a separate selection kernel cannot share an SM with a pass that holds every register, so
    even on a parallel graph branch it delayed the column pass).
    ``zero_grad``: the caller knows that a == b bit for bit (first inner iteration of an SVRG epoch, right after
    ``w = copy(z)``), so g_sel(a - b) is exactly zero whatever the minibatch: on the fused CSMRI path the forward line
    pass, the column pass and the inverse transforms are skipped (the tail kernel runs on a zero spectrum; same bits as
    transforming zeros), the minibatch counters still advance.  Other paths ignore the hint.
    ``next_pass`` = dict(idx=positions of the NEXT minibatch or None for the device sampler): the tail kernel also runs the
    forward line pass of the next iteration (z_new - b) and builds its selection (pnp_csmri_update_prox_next); the call for
    that iteration then says ``r2c_done=True`` and starts at the column pass.  Only honoured on the fused CSMRI path
    (``eng.fused_tail is True``); whole-epoch graphs use it (SvrgRun._epoch_ops)."""
    p, d = eng.p, eng.d
    own = sel is not None and sel is eng.sel
    if (eng.fused_tail is not False and eng.uses_sigma and not eng.sigma_ready and getattr(d, 'method', None) == 'wavelet'
            and hasattr(p, '_dev_update_prox')):
        kw = dict(b=b, sel=sel, with_y=False, gscale=gscale, vadd=vadd, step_ptr=eng.step, z_in=z, z_out=z, clear_sel=own)
        skip = bool(zero_grad) and eng.fused_tail is True and own          # (the first fused call decides whether the tail applies)
        if eng.chain:
            kw['chain'] = True
        if skip:
            pass
        elif r2c_done and eng.fused_tail is True and own:
            p._dev_grad(a, phases=2, **kw)          # spectrum of a - b and the selection were left by the previous tail launch
        else:
            job = eng.sel_job() if (sel_fn is not None and own) else None
            if job is None and sel_fn is not None:
                sel_fn()
            p._dev_grad(a, phases=3, sel_job=job, **kw)
        nxt = None
        if next_pass is not None and eng.fused_tail is True and own:
            job = eng.sel_job()
            if job is not None:
                job = dict(job, counter_add=1)
                if job.get('idx') is not None:
                    job['idx'] = next_pass.get('idx')
            nxt = dict(w=b, sel=sel, job=job)
        ok = p._dev_update_prox(gscale, eng.step, vadd, z, z, eng.sig_log, d.sigma_modifier,
                                d.denoise_strength * d.decay ** (d.t + 1), p._xrec_dev, eng.mse_log, eng.slot_ptr,
                                advance=eng.counters if advance else None, n_advance=advance,
                                barrier_ws=eng.barrier_ws if (eng.chain or eng.sw_barrier) else None, chain=eng.chain,
                                zero_spectrum=skip, next_pass=nxt)
        eng.fused_tail = ok
        if ok:
            d.t += 1
            return
        p._dev_grad(a, phases=4, **kw)              # the spectrum is there: finish with the separate kernels
    else:
        if sel_fn is not None:
            sel_fn()
        _grad_update(eng, a, b, sel, False, gscale, vadd=vadd, step_ptr=eng.step, z_in=z, z_out=z)
    eng.prox(z, z)
    if advance:
        eng.advance(advance)


class _Faithful:
    """One eager iteration with the reference's wall-clock phase timing and PSNR read-back."""

    def __init__(self, eng, verbose):
        self.eng, self.verbose = eng, verbose

    def run(self, grad_phase, prox_in, prox_out, label_before=None, label_after=None, gd_style_time=False):
        eng = self.eng
        t0 = time.time()
        with torch.cuda.stream(eng.stream):
            grad_phase()
        eng.stream.synchronize()
        g_t = time.time() - t0
        eng.gradient_time += g_t
        if self.verbose and label_before is not None:
            print(label_before + str(eng.psnr_of(prox_in)))
        t1 = time.time()
        with torch.cuda.stream(eng.stream):
            eng.prox(prox_in, prox_out)
            eng.advance()
            psnr = eng.read_slot()
        d_t = time.time() - t1
        eng.denoise_time += d_t
        eng.time_log.append(time.time() - t0 if gd_style_time else g_t + d_t)
        eng.psnr_log.append(psnr)
        if self.verbose and label_after is not None:
            print(label_after + str(psnr))
        return psnr


def _fast_inner(eng, budget, n_iters, inner_ops, host_draw, sync_every, converge_check, diverge_check, start_psnr):
    """Run up to n_iters graph-replayed inner iterations; returns (psnr_z, stop, done_iters)."""
    # minibatches: device sampler draws in-graph; host sources are pre-drawn into a device ring
    if eng.graph is None:
        eng.graph = eng.capture(inner_ops)
    done = 0
    stop = False
    psnr_z = start_psnr
    if not (converge_check is True or diverge_check is True):
        # no stop rule needs the PSNR of every iterate: enqueue, count, and read the log back later
        while done < n_iters and budget.alive():
            left = budget.left()
            # never enqueue more than sync_every iterations between two looks at the wall clock (``tt`` is the
            # reference's only loop bound, algorithms/pnp_svrg.py:28; tune_* trials rely on it)
            room = min(n_iters - done, LOG_CHUNK - eng.slot_host, max(1, sync_every - eng.since_sync),
                       left if left is not None else 1 << 30)
            if room <= 0:
                eng.resolve()
                continue
            with torch.cuda.stream(eng.stream):
                for _ in range(room):
                    if host_draw is not None:
                        host_draw()
                    eng.replay(eng.graph)
            eng.defer_slots(room)
            budget.calls += room
            done += room
            if eng.since_sync >= sync_every:
                eng.resolve()              # bounds how far the host runs ahead of the GPU (wall-clock budget tt)
        return None, False, done
    while done < n_iters and not stop:
        left = budget.left()
        room = min(n_iters - done, sync_every, LOG_CHUNK - eng.slot_host, left if left is not None else 1 << 30)
        if room <= 0:
            break
        t0 = time.time()
        with torch.cuda.stream(eng.stream):
            for _ in range(room):
                if host_draw is not None:
                    host_draw()
                eng.replay(eng.graph)
        eng.slot_host += room
        vals = eng.flush_fast()
        dt = time.time() - t0
        for v in vals:
            eng.time_log.append(dt / len(vals))
            eng.psnr_log.append(v)
            budget.calls += 1
            eng.n_prox += 1
            done += 1
            if stop_rule(psnr_z, v, converge_check, diverge_check):
                stop = True
            psnr_z = v
        if not budget.alive():
            break
    return psnr_z, stop, done


def _host_draw_fn(eng, extra_fn=None):
    if eng.B <= 0 or eng.mb_source == 'device':
        return None

    if eng.mb_source == 'host':
        # look-ahead draws: one native call waits for the draw, adds the extras (independent generator, so the order
        # against the minibatch draw does not matter) and enqueues the pinned -> device copy
        def draw_native():
            eng.draw_host(extra_fn() if extra_fn else ())
        return draw_native

    def draw():
        # a small ring of pinned staging buffers: the host only waits for the H2D copy that used this buffer four
        # draws ago, so drawing and staging overlap the GPU's previous iterations instead of a sync per iteration
        eng.next_host_buffer()
        eng.draw_host()                                   # minibatch first, then the extras
        if extra_fn:                                      # (reference RNG call order, pnp_saga.py:43-44)
            for k, e in enumerate(extra_fn()):
                eng.idx_host.numpy()[eng.B + k] = e
        eng.idx_dev.copy_(eng.idx_host, non_blocking=True)
        eng.mark_host_buffer()
    return draw


def _sel_ops(eng):
    """selection rebuild inside the (possibly captured) iteration"""
    if eng.mb_source == 'device':
        eng.sample_sel_device()
    else:
        eng.p._dev_set_sel(eng.sel, eng.idx_dev, eng.B, clear=False)


# ------------------------------------------------------------------------------------------ GD
def pnp_gd(problem, denoiser, eta, tt, verbose=True, lr_decay=1, converge_check=True, diverge_check=False,
           max_iters=None, fast=False, sync_every=64):
    """algorithms/pnp_gd.py:8-84."""
    eng = Engine(problem, denoiser, fast=fast)
    budget = Budget(tt, max_iters)
    eng.time_log.append(time.time() - budget.t0)
    psnr_z = eng.psnr_of(eng.z)
    eng.psnr_log.append(psnr_z)
    z = eng.z
    i = 0
    M0 = getattr(problem, 'M0', problem.M)

    def grad_phase():
        _grad_update(eng, z, None, None, True, 1.0 / _full_norm(problem), step_ptr=eng.step, z_in=z, z_out=z)

    if fast:
        with torch.cuda.stream(eng.stream):
            eng.set_step(eta)

        def ops():
            grad_phase()
            eng.prox(z, z)
            eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 3, D.ptr(eng.step), float(lr_decay), eng.sptr))
        while budget.alive():
            psnr_z, stop, done = _fast_inner(eng, budget, 1 << 30, ops, None, sync_every, converge_check,
                                             diverge_check, psnr_z)
            if stop or done == 0:
                break
        eng.destroy(eng.graph)
        return eng.result('PnP GD')

    it = _Faithful(eng, verbose)
    while budget.alive():
        start = psnr_z
        with torch.cuda.stream(eng.stream):
            eng.set_step(eta * lr_decay ** i)
        psnr_z = it.run(grad_phase, z, z, str(i) + " Before denoising:  ", str(i) + " After denoising:  ",
                        gd_style_time=True)
        budget.calls += 1
        i += 1
        if stop_rule(start, psnr_z, converge_check, diverge_check):
            break
    return eng.result('PnP GD')


def _full_norm(problem):
    """divisor of grad_full: M0 for CSMRI (problems/CSMRI.py:81), M otherwise (DeblurSR.py:132, PR.py:79)"""
    return problem.M0 if getattr(problem, 'pname', '') == 'csmri' else problem.M


# ----------------------------------------------------------------------------------------- SGD
def pnp_sgd(problem, denoiser, eta, tt, mini_batch_size, verbose=True, lr_decay=1, converge_check=True,
            diverge_check=False, max_iters=None, mb_source='legacy', mb_seed=0, mb_stream=None, fast=False,
            sync_every=64):
    """algorithms/pnp_sgd.py:8-84."""
    B = int(mini_batch_size)
    eng = Engine(problem, denoiser, B, mb_source, mb_seed, mb_stream, fast)
    budget = Budget(tt, max_iters)
    eng.time_log.append(time.time() - budget.t0)
    psnr_z = eng.psnr_of(eng.z)
    eng.psnr_log.append(psnr_z)
    z = eng.z
    i = 0

    def grad_ops():
        _grad_update(eng, z, None, eng.sel, True, 1.0 / B, step_ptr=eng.step, z_in=z, z_out=z)

    if fast:
        with torch.cuda.stream(eng.stream):
            eng.set_step(eta)

        def ops():
            _sel_ops(eng)
            grad_ops()
            eng.prox(z, z)
            eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 3, D.ptr(eng.step), float(lr_decay), eng.sptr))
        draw = _host_draw_fn(eng)
        while budget.alive():
            psnr_z, stop, done = _fast_inner(eng, budget, 1 << 30, ops, draw, sync_every, converge_check,
                                             diverge_check, psnr_z)
            if stop or done == 0:
                break
        eng.destroy(eng.graph)
        return eng.result('PnP SGD')

    it = _Faithful(eng, verbose)
    while budget.alive():
        start = psnr_z

        def grad_phase():
            eng.set_step(eta * lr_decay ** i)
            if mb_source == 'device':
                eng.sample_sel_device()
            else:
                eng.draw_host()
                eng.upload_sel()
            grad_ops()
        psnr_z = it.run(grad_phase, z, z, str(i) + " Before denoising:  ", str(i) + " After denoising:  ")
        budget.calls += 1
        i += 1
        if stop_rule(start, psnr_z, converge_check, diverge_check):
            break
    return eng.result('PnP SGD')


# ---------------------------------------------------------------------------------------- SVRG
class SvrgRun:
    """One PnP-SVRG reconstruction resident on the device (algorithms/pnp_svrg.py:8-105): the iterate, the snapshot
    ``w``, the snapshot gradient ``mu`` and the logs live in HBM, ``loop()`` is the reference's control flow.

    ``fast`` without stop rules runs WHOLE EPOCHS as one CUDA graph each (``epoch()``): snapshot gradient, ``w = z``,
    T2 inner iterations (minibatch selection on a parallel branch next to the forward line pass, fused gradient +
    variance-reduced update + sigma estimate + prox + PSNR), step decay.  Host-drawn minibatches ('host', 'legacy',
    'stream') of epoch e + 1 are staged and copied on a second stream while the graph of epoch e runs: two sets of
    T2 device index buffers, one graph per set.  ``bench.py`` times exactly this method; ``pnp_svrg(fast=True)`` is a
    loop around it."""

    def __init__(self, problem, denoiser, eta, T2, mini_batch_size, lr_decay=1, vr_mode='as_committed',
                 mb_source='legacy', mb_seed=0, mb_stream=None, fast=False):
        if vr_mode not in ('as_committed', 'paper'):
            raise ValueError("vr_mode must be 'as_committed' or 'paper'")
        self.problem, self.denoiser = problem, denoiser
        self.eta, self.lr_decay = float(eta), float(lr_decay)
        self.B, self.T2 = int(mini_batch_size), int(T2)
        self.paper = vr_mode == 'paper'
        self.mb_source, self.fast = mb_source, fast
        self.eng = eng = Engine(problem, denoiser, self.B, mb_source, mb_seed, mb_stream, fast)
        self.z = eng.z
        with torch.cuda.stream(eng.stream):
            self.w = torch.empty_like(self.z)
            self.mu = torch.empty_like(self.z)
        self.i = 0                      # outer (epoch) index
        self.draw = _host_draw_fn(eng) if (self.paper or mb_source == 'legacy') else None
        self._epoch_graphs = {}
        self._inflight = []             # one event per epoch in flight
        self.max_ahead = 4              # epochs the host may run ahead of the GPU
        self._epoch_sets = None         # device index buffers [set][T2][B] for host-drawn minibatches
        self._step_on_device = False    # epoch graphs keep eta * lr_decay**i on the device

    # ---- the launches ----------------------------------------------------------------------------------------
    def snapshot(self):
        """mu = grad_full(z) ; w = copy(z)          (pnp_svrg.py:32-35)"""
        eng, problem = self.eng, self.problem
        if getattr(problem, 'shard', None) is not None:
            # measurement-sharded snapshot: every rank transforms its band of packed ky rows, the partial gradients are
            # summed by one NCCL all-reduce of 4N bytes
            _grad_update(eng, self.z, None, None, True, 1.0 / _full_norm(problem), g_out=self.mu, partial_ok=True)
            problem._snapshot_allreduce(self.mu)
        else:
            _grad_update(eng, self.z, None, None, True, 1.0 / _full_norm(problem), g_out=self.mu)
        eng.copy(self.w, self.z)

    def grad_ops(self):
        eng, z = self.eng, self.z
        if self.paper:
            _grad_update(eng, z, self.w, eng.sel, False, 1.0 / self.B, vadd=self.mu, step_ptr=eng.step, z_in=z, z_out=z)
        else:
            eng.check(eng.lib.pnp_axpy(D.ptr(z), D.ptr(self.mu), D.ptr(z), eng.N, 1, 0.0, D.ptr(eng.step), eng.sptr))

    def fast_ops(self, first_of_epoch=False, r2c_done=False, next_pass=None):
        """one inner iteration as it is captured (per iteration, or T2 times inside an epoch graph).
        ``first_of_epoch``: z == w (the snapshot has just copied it), the stochastic term of line 53 is exactly zero;
        ``r2c_done`` / ``next_pass``: see _grad_update_prox (forward line pass fused into the previous tail launch)"""
        eng, z = self.eng, self.z
        if self.paper:
            _grad_update_prox(eng, z, self.w, eng.sel, 1.0 / self.B, self.mu, z, advance=3,
                              sel_fn=lambda: _sel_ops(eng), zero_grad=first_of_epoch, r2c_done=r2c_done, next_pass=next_pass)
        else:
            self.grad_ops()
            eng.prox(z, z)
            eng.advance()

    # ---- whole epochs as one graph ----------------------------------------------------------------------------
    def epoch_mode(self, converge_check=False, diverge_check=False):
        """True when whole epochs can run as one graph: fast mode, no per-iterate stop rule, no sharded snapshot
        (its NCCL all-reduce stays outside a capture), log capacity for an epoch."""
        return (self.fast and not (converge_check is True or diverge_check is True) and self.T2 >= 1
                and getattr(self.problem, 'shard', None) is None and self.T2 <= LOG_CHUNK // 4)

    def _needs_indices(self):
        return self.B > 0 and self.mb_source != 'device' and (self.paper or self.mb_source == 'legacy')

    def _small_ok(self):
        """True when a whole epoch runs as ONE launch of the cluster kernel (pnp_csmri_svrg_small, csrc/small.cuh):
        paper-mode PnP-SVRG on a square CSMRI image of 128 or 256 pixels a side with the wavelet prox.
        ``PNP_SMALL=0`` keeps the three-pass path (comparison runs, tests)."""
        p, d, eng = self.problem, self.denoiser, self.eng
        return (self.paper and self.B > 0 and os.environ.get('PNP_SMALL', '1') != '0'
                and getattr(p, 'pname', '') == 'csmri' and getattr(p, 'shard', None) is None
                and getattr(d, 'method', None) == 'wavelet' and eng.uses_sigma and hasattr(p, '_bits_full')
                and eng.lib.pnp_csmri_svrg_small_supported(int(p.H), int(p.W)) == 1)

    def _small_epoch(self, bufs):
        """snapshot + T2 inner iterations + counters (+ step decay) with the image resident in a cluster's shared memory"""
        import ctypes as C
        from .. import _lib
        eng, p, d = self.eng, self.problem, self.denoiser
        idx, stride = None, 0
        if bufs is not None:                    # host-drawn minibatches: the T2 index sets of this epoch, contiguous
            idx, stride = bufs[0], int(bufs[0].numel())
        args = _lib.SvrgSmallArgs(
            H=p.H, W=p.W, batch=1, z=D.ptr(self.z), xrec=D.ptr(p._xrec_dev),
            Y1=D.ptr(p._Y1), Y2=D.ptr(p._Y2), Y1n=D.ptr(p._Y1n), Y2n=D.ptr(p._Y2n), bits_full=D.ptr(p._bits_full),
            support=D.ptr(p._support), m0=D.ptr(p._m0_dev), support_img_stride=0,
            idx=D.ptr(idx), idx_img_stride=0, idx_iter_stride=stride,
            snap_scale_ptr=None, snap_scale=1.0 / _full_norm(p), step=D.ptr(eng.step), step_img_stride=0,
            sig_log=D.ptr(eng.sig_log), mse_log=D.ptr(eng.mse_log), slot=D.ptr(eng.slot_ptr), draw_counter=D.ptr(eng.draw_ptr),
            n_inner=self.T2, T2=self.T2, mini_batch_size=self.B, seed=eng.mb_seed & 0xffffffff, lr_decay=1.0,
            sigma_modifier=float(d.sigma_modifier), fallback_sigma=float(d.denoise_strength * d.decay ** (d.t + 1)),
            fallback_decay=float(d.decay))
        eng.check(eng.lib.pnp_csmri_svrg_small(C.byref(args), eng.sptr))
        eng.check(eng.lib.pnp_advance_by(D.ptr(eng.counters), 3, self.T2, eng.sptr))
        if self.lr_decay != 1.0:
            eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 0, D.ptr(eng.step), self.lr_decay, eng.sptr))

    def _epoch_ops(self, bufs):
        eng = self.eng
        if self._small_ok():
            self._small_epoch(bufs)
            return
        self.snapshot()
        keep = eng.idx_dev if self.B > 0 else None
        # CSMRI + wavelet prox, PNP_FUSE_R2C=1: the tail launch of iteration j also runs the forward line pass and the
        # selection of iteration j + 1 (pnp_csmri_update_prox_next).  Bit-identical and tested, but OFF by default: measured
        # on B200 at 2048^2 it is 63.7 us per inner iteration against 62.6-63.8 with the separate pass -- the two fused
        # rounds of transforms take 11.4 us inside the 16-warp cooperative kernel (snapshot lines from DRAM without a TMA
        # double buffer, cold instruction cache), as long as the stand-alone pass with its launch and cold start.
        fuse = self.paper and eng.fused_tail is True and os.environ.get('PNP_FUSE_R2C', '0') == '1'
        for j in range(self.T2):
            if bufs is not None:
                eng.idx_dev = bufs[j]
            nxt = None
            if fuse and j + 1 < self.T2:
                nxt = dict(idx=bufs[j + 1] if bufs is not None else None)
            self.fast_ops(first_of_epoch=(j == 0), r2c_done=(fuse and j >= 1), next_pass=nxt)
        if self.B > 0:
            eng.idx_dev = keep
        if self.lr_decay != 1.0:
            eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 0, D.ptr(eng.step), self.lr_decay, eng.sptr))

    def _prepare_epochs(self):
        eng = self.eng
        if self._epoch_graphs:
            return
        n_sets = 2 if self._needs_indices() else 1
        if n_sets == 2:
            n = self.B + eng.n_extra
            with torch.cuda.stream(eng.stream):
                self._epoch_sets = torch.zeros(2, self.T2, n, dtype=torch.int32, device=eng.dev)
            self._copy_stream = torch.cuda.Stream(device=eng.dev)
            self._ev_ready = [torch.cuda.Event() for _ in range(2)]
            self._ev_done = [None, None]
            if eng.mb_source != 'host':            # 'legacy' / 'stream': drawn in Python into pinned sets
                self._pinned_sets = torch.empty(2, self.T2, n, dtype=torch.int32).pin_memory()
                self._ev_copied = [None, None]
        with torch.cuda.stream(eng.stream):
            eng.set_step(self.eta * self.lr_decay ** self.i)
            # one eager epoch-shaped warm-up is NOT run: the kernels were loaded by pnp_init, and a capture does not execute
        if eng.fused_tail is None and hasattr(self.problem, '_dev_update_prox'):
            # known before the capture, so that the first inner iteration of every captured epoch can skip its transforms
            eng.fused_tail = True if eng.lib.pnp_csmri_update_prox_supported(int(self.problem.H), int(self.problem.W)) == 1 else None
        t_before = self.denoiser.t
        # inside the epoch graph the passes form one chain on one stream: programmatic dependent launch lets every
        # kernel start its prologue (and the loads that do not depend on its predecessor) while the previous one drains
        # (measured on B200, profiles/README.md: neither it nor the software grid barrier it needs in the tail kernel
        # gains anything over plain / cooperative launches -- the cost of a kernel boundary here is the drain and fill
        # of the persistent grids, not the launch latency -- so both stay off unless PNP_CHAIN=1 / PNP_SW_BARRIER=1)
        eng.chain = bool(getattr(self.problem, '_chain_ok', False)) and os.environ.get('PNP_CHAIN', '0') == '1'
        eng.sw_barrier = bool(getattr(self.problem, '_chain_ok', False)) and os.environ.get('PNP_SW_BARRIER', '0') == '1'
        try:
            for s_ in range(n_sets):
                bufs = None if self._epoch_sets is None else [self._epoch_sets[s_, j] for j in range(self.T2)]
                self._epoch_graphs[s_] = eng.capture(lambda: self._epoch_ops(bufs))
        finally:
            eng.chain = False
            eng.sw_barrier = False
        self.denoiser.t = t_before      # a capture runs the host side of fast_ops only
        self._step_on_device = True
        self._epochs_launched = 0

    def _stage_epoch(self, s_):
        """draw the T2 minibatches of the next epoch and copy them into index set ``s_`` on the copy stream"""
        eng, cs = self.eng, self._copy_stream
        if self._ev_done[s_] is not None:
            cs.wait_event(self._ev_done[s_])           # the epoch that last read this set (two epochs ago)
        if eng.mb_source == 'host':
            # native look-ahead queue: draw -> pinned ring -> async copy
            if os.environ.get('PNP_STAGE_MANY', '1') != '0':
                # the T2 draws of the epoch in one call
                eng._draws.stage_many(self._epoch_sets[s_].data_ptr(), self.T2, int(self._epoch_sets.shape[2]), cs.cuda_stream)
            else:
                for j in range(self.T2):
                    eng._draws.stage(self._epoch_sets[s_, j].data_ptr(), (), cs.cuda_stream)
        else:
            if self._ev_copied[s_] is not None:
                self._ev_copied[s_].synchronize()      # the copy that last read this pinned set
            host = self._pinned_sets[s_].numpy()
            for j in range(self.T2):
                if eng.mb_source == 'legacy':
                    host[j, :self.B] = eng.p._draw_indices(self.B)      # np.random, the reference's call order
                else:
                    idx = np.asarray(eng.mb_stream[eng._stream_pos])
                    eng._stream_pos += 1
                    if idx.size != self.B:
                        raise ValueError('mb_stream entry %d has %d positions, expected %d' % (eng._stream_pos - 1, idx.size, self.B))
                    host[j, :self.B] = idx
            with torch.cuda.stream(cs):
                self._epoch_sets[s_].copy_(self._pinned_sets[s_], non_blocking=True)
            ev = self._ev_copied[s_] or torch.cuda.Event()
            ev.record(cs)
            self._ev_copied[s_] = ev
        self._ev_ready[s_].record(cs)

    def epoch(self):
        """Enqueue one whole epoch (snapshot + T2 inner iterations); nothing is read back (``eng.defer_slots``)."""
        eng = self.eng
        self._prepare_epochs()
        if eng.slot_host + self.T2 > LOG_CHUNK:
            eng.resolve()
        s_ = 0
        if self._epoch_sets is not None:
            s_ = self._epochs_launched & 1
            self._stage_epoch(s_)
            eng.stream.wait_event(self._ev_ready[s_])
        if eng.deferred:
            eng.defer_dup()                              # PnP-SVRG logs PSNR(z) again at every snapshot (pnp_svrg.py:37-40)
        else:
            eng.time_log.append(0.0)
            eng.psnr_log.append(eng.psnr_log[-1])
        eng.replay(self._epoch_graphs[s_])
        if self._epoch_sets is not None:
            ev = self._ev_done[s_] or torch.cuda.Event()
            ev.record(eng.stream)
            self._ev_done[s_] = ev
        # bound how far the host runs ahead of the GPU without draining the stream: wait for the epoch launched
        # `max_ahead` epochs ago (the wall-clock budget tt is looked at between epochs)
        ring = self._inflight
        ev = torch.cuda.Event()
        ev.record(eng.stream)
        ring.append(ev)
        if len(ring) > self.max_ahead:
            ring.pop(0).synchronize()
        eng.defer_slots(self.T2)
        self.denoiser.t += self.T2
        self._epochs_launched += 1
        self.i += 1

    def close(self):
        eng = self.eng
        eng.destroy(eng.graph)
        eng.graph = None
        for g in self._epoch_graphs.values():
            eng.destroy(g)
        self._epoch_graphs = {}

    # ---- the reference's loop ---------------------------------------------------------------------------------
    def loop(self, tt, max_iters=None, verbose=True, converge_check=True, diverge_check=False, sync_every=64):
        eng, z, T2, fast = self.eng, self.z, self.T2, self.fast
        budget = Budget(tt, max_iters)
        eng.time_log.append(time.time() - budget.t0)
        psnr_z = eng.psnr_of(z)
        eng.psnr_log.append(psnr_z)
        stop = False
        it = _Faithful(eng, verbose)
        epochs = self.epoch_mode(converge_check, diverge_check)
        while budget.alive() and not stop:
            left = budget.left()
            if epochs and (left is None or left >= T2):
                self.max_ahead = max(1, min(8, int(sync_every) // T2))    # run-ahead bound, in epochs (wall-clock budget tt)
                self.epoch()
                budget.calls += T2
                psnr_z = None
                continue
            t_outer = time.time()
            with torch.cuda.stream(eng.stream):
                self.snapshot()
                if not self._step_on_device:
                    eng.set_step(self.eta * self.lr_decay ** self.i)
            if not fast:
                eng.stream.synchronize()
            if eng.deferred:
                eng.defer_dup()
            else:
                eng.time_log.append(time.time() - t_outer)
                eng.psnr_log.append(psnr_z if psnr_z is not None else eng.psnr_log[-1])
            if fast:
                psnr_z, stop, _ = _fast_inner(eng, budget, T2, self.fast_ops, self.draw,
                                              min(sync_every, T2) if converge_check or diverge_check else sync_every,
                                              converge_check, diverge_check, psnr_z)
                if self._step_on_device and self.lr_decay != 1.0:
                    with torch.cuda.stream(eng.stream):
                        eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 0, D.ptr(eng.step), self.lr_decay, eng.sptr))
            else:
                i = self.i
                for j in range(T2):
                    if not budget.alive():
                        break
                    start = psnr_z

                    def grad_phase():
                        if self.mb_source == 'device':
                            if self.paper:
                                eng.sample_sel_device()
                        else:
                            eng.draw_host()            # drawn even when unused (pnp_svrg.py:52)
                            if self.paper:
                                eng.upload_sel()
                        self.grad_ops()
                    psnr_z = it.run(grad_phase, z, z, str(i) + " " + str(j) + " Before denoising:  ",
                                    "After denoising update: " + str(i) + " " + str(j) + " ")
                    budget.calls += 1
                    if stop_rule(start, psnr_z, converge_check, diverge_check):
                        stop = True
                        break
            self.i += 1
        self.close()
        return eng.result('PnP SVRG')


def pnp_svrg(problem, denoiser, eta, tt, T2, mini_batch_size, verbose=True, lr_decay=1, converge_check=True,
             diverge_check=False, max_iters=None, vr_mode='as_committed', mb_source='legacy', mb_seed=0,
             mb_stream=None, fast=False, sync_every=64):
    """algorithms/pnp_svrg.py:8-105."""
    run = SvrgRun(problem, denoiser, eta, T2, mini_batch_size, lr_decay, vr_mode, mb_source, mb_seed, mb_stream, fast)
    return run.loop(tt, max_iters, verbose, converge_check, diverge_check, sync_every)


# ---------------------------------------------------------------------------------------- SAGA
def pnp_saga(problem, denoiser, eta, tt, mini_batch_size, hist_size=50, verbose=True, lr_decay=1,
             converge_check=True, diverge_check=False, max_iters=None, mb_source='legacy', mb_seed=0,
             mb_stream=None, fast=False, sync_every=64):
    """algorithms/pnp_saga.py:8-102 (including its non-textbook use of the previous iteration's
    gradient, :47,72)."""
    B = int(mini_batch_size)
    hist = int(hist_size)
    eng = Engine(problem, denoiser, B, mb_source, mb_seed, mb_stream, fast, n_extra_ints=1)
    budget = Budget(tt, max_iters)
    z = eng.z
    slot_rng = np.random.default_rng(mb_seed + 1) if mb_source != 'legacy' else None
    with torch.cuda.stream(eng.stream):
        g_new = torch.empty_like(z)
        g_prev = torch.empty_like(z)
        table = torch.empty(hist * eng.N, dtype=torch.float32, device=eng.dev)
        tsum = torch.empty_like(z)
    slot_dev = eng.idx_dev[B:B + 1]

    def draw_slot():
        if slot_rng is None:
            return (int(np.random.choice(hist, 1).item()),)     # pnp_saga.py:44
        return (int(slot_rng.integers(hist)),)

    t0 = time.time()
    with torch.cuda.stream(eng.stream):
        # stoch_init / table fill                       (pnp_saga.py:25-29)
        if mb_source == 'device':
            eng.sample_sel_device()
            eng.advance()
        else:
            eng.draw_host()
            eng.upload_sel()
        _grad_update(eng, z, None, eng.sel, True, 1.0 / B, g_out=g_prev)
        eng.check(eng.lib.pnp_saga_init(D.ptr(g_prev), D.ptr(table), D.ptr(tsum), eng.N, 1, hist, eng.sptr))
    eng.stream.synchronize()
    if mb_source == 'device':
        with torch.cuda.stream(eng.stream):
            eng.counters[0:2].zero_()
    eng.time_log.append(time.time() - t0)
    psnr_z = eng.psnr_of(z)
    eng.psnr_log.append(psnr_z)

    def grad_ops():
        _grad_update(eng, z, None, eng.sel, True, 1.0 / B, g_out=g_new)
        eng.check(eng.lib.pnp_saga_update(D.ptr(g_new), D.ptr(g_prev), D.ptr(table), D.ptr(tsum), D.ptr(z), eng.N, 1,
                                          hist, D.ptr(slot_dev), 0, None, 0.0, D.ptr(eng.step), eng.sptr))

    i = 0
    if fast:
        with torch.cuda.stream(eng.stream):
            eng.set_step(eta)

        def ops():
            _sel_ops(eng)
            grad_ops()
            eng.prox(z, z)
            eng.check(eng.lib.pnp_advance_scale(D.ptr(eng.counters), 3, D.ptr(eng.step), float(lr_decay), eng.sptr))
        if mb_source == 'device':
            def draw():
                eng.idx_host.numpy()[B] = draw_slot()[0]
                eng.idx_dev[B:B + 1].copy_(eng.idx_host[B:B + 1], non_blocking=True)
                eng.stream.synchronize()
        else:
            draw = _host_draw_fn(eng, draw_slot)
        while budget.alive():
            psnr_z, stop, done = _fast_inner(eng, budget, 1 << 30, ops, draw, sync_every, converge_check,
                                             diverge_check, psnr_z)
            if stop or done == 0:
                break
        eng.destroy(eng.graph)
        return eng.result('pnp_saga')

    it = _Faithful(eng, verbose)
    while budget.alive():
        start = psnr_z

        def grad_phase():
            eng.set_step(eta * lr_decay ** i)
            if mb_source == 'device':
                eng.sample_sel_device()
                eng.idx_host.numpy()[B] = draw_slot()[0]
                eng.idx_dev[B:B + 1].copy_(eng.idx_host[B:B + 1], non_blocking=True)
            else:
                # reference order: select_mb, then the table slot (pnp_saga.py:43-44)
                idx = eng.p._draw_indices(B) if mb_source == 'legacy' else None
                if idx is None:
                    eng.draw_host(draw_slot())
                else:
                    buf = eng.idx_host.numpy()
                    buf[:B] = idx
                    buf[B] = draw_slot()[0]
                eng.upload_sel()
            grad_ops()
        psnr_z = it.run(grad_phase, z, z, str(i) + " Before denoising:  ", str(i) + " After denoising:  ")
        budget.calls += 1
        i += 1
        if stop_rule(start, psnr_z, converge_check, diverge_check):
            break
    return eng.result('pnp_saga')


# --------------------------------------------------------------------------------------- SARAH
def pnp_sarah(problem, denoiser, eta, tt, T2, mini_batch_size, verbose=True, lr_decay=1, converge_check=True,
              diverge_check=False, max_iters=None, mb_source='legacy', mb_seed=0, mb_stream=None, fast=False,
              sync_every=64):
    """algorithms/pnp_sarah.py:8-129 (w_next is never advanced inside the inner loop and z is not
    reset to it -- kept as in the reference, :60-104)."""
    B = int(mini_batch_size)
    T2 = int(T2)
    eng = Engine(problem, denoiser, B, mb_source, mb_seed, mb_stream, fast)
    budget = Budget(tt, max_iters)
    z = eng.z
    with torch.cuda.stream(eng.stream):
        w_prev = torch.empty_like(z)
        w_next = torch.empty_like(z)
        v_prev = torch.empty_like(z)
    psnr_z = eng.psnr_of(z)            # start_PSNR of the first inner iteration (not logged, :65)
    it = _Faithful(eng, verbose)

    def outer_grad():
        # w_prev = z ; v_prev = grad_full(z) ; w_next = w_prev - eta * v_prev   (no lr_decay, :28-36)
        eng.copy(w_prev, z)
        _grad_update(eng, z, None, None, True, 1.0 / _full_norm(problem), g_out=v_prev, step=float(eta),
                     z_in=z, z_out=w_next)

    def grad_ops():
        # v_next = (g_B(w_next) - g_B(w_prev)) / B + v_prev ; z -= step * v_next ; v_prev = v_next  (:72-75,97)
        _grad_update(eng, w_next, w_prev, eng.sel, False, 1.0 / B, vadd=v_prev, v_out=v_prev, step_ptr=eng.step,
                     z_in=z, z_out=z)

    def after_prox():
        eng.copy(w_prev, z)                     # w_previous = z0   (:98)

    def fast_ops():
        _sel_ops(eng)
        grad_ops()
        eng.prox(z, z)
        after_prox()
        eng.advance()
    draw = _host_draw_fn(eng)

    i = 0
    stop = False
    while budget.alive() and not stop:
        # the outer prox step is logged like an iteration (:38-50)
        eng.resolve()
        t0 = time.time()
        with torch.cuda.stream(eng.stream):
            outer_grad()
        eng.stream.synchronize()
        g_t = time.time() - t0
        eng.gradient_time += g_t
        t1 = time.time()
        with torch.cuda.stream(eng.stream):
            eng.prox(w_next, w_next)
            eng.advance()
            psnr_w = eng.read_slot()
        d_t = time.time() - t1
        eng.denoise_time += d_t
        eng.time_log.append(g_t + d_t)
        eng.psnr_log.append(psnr_w)
        budget.calls += 1
        with torch.cuda.stream(eng.stream):
            eng.set_step(eta * lr_decay ** i)
        if fast:
            psnr_z, stop, _ = _fast_inner(eng, budget, T2, fast_ops, draw, min(sync_every, T2) if converge_check or diverge_check else sync_every,
                                          converge_check, diverge_check, psnr_z)
        else:
            for j in range(T2):
                if not budget.alive():
                    break
                start = psnr_z

                def grad_phase():
                    if mb_source == 'device':
                        eng.sample_sel_device()
                    else:
                        eng.draw_host()
                        eng.upload_sel()
                    grad_ops()
                psnr_z = it.run(grad_phase, z, z, "After gradient update: " + str(i) + " " + str(j) + " ",
                                "After denoising update: " + str(i) + " " + str(j) + " ")
                with torch.cuda.stream(eng.stream):
                    after_prox()
                budget.calls += 1
                if stop_rule(start, psnr_z, converge_check, diverge_check):
                    stop = True
                    break
        i += 1
    eng.destroy(eng.graph)
    return eng.result('pnp_sarah')
