"""Device-resident PnP iteration engine shared by algorithms/pnp_{gd,sgd,svrg,saga,sarah}.

The reference loops (algorithms/pnp_*.py) keep the iterate in host NumPy and call
problem.grad_* / estimate_sigma / denoiser.denoise / problem.PSNR every iteration.  Here the
iterate, the snapshot, the gradient tables and the PSNR / sigma logs stay in HBM; one inner
iteration is a fixed sequence of kernel launches on one CUDA stream:

    [selection bits] -> [gradient + variance-reduced update] -> [sigma estimate] -> [prox + PSNR]
    -> [advance device counters]

Two ways of driving that sequence:
  faithful (default)  launches are issued eagerly, the minibatch is drawn on the host with
                      NumPy's legacy global RNG in the reference's call order, and the PSNR of
                      every iterate is read back (8 bytes) so converge_check / diverge_check /
                      verbose behave exactly like the reference;
  fast                the sequence is captured once in a CUDA graph and replayed; minibatches
                      come from a pre-drawn device buffer or the device sampler, logs are read
                      back once per chunk, stop rules are applied at chunk boundaries.
"""
import os
import time

import numpy as np
import torch

from . import _lib, device as D

TOL = 1e-5
LOG_CHUNK = 4096


class ProxCtx:
    """What a denoiser's _dev_denoise needs."""
    __slots__ = ('z_in', 'z_out', 'H', 'W', 'sig_log', 'sigma_est', 'xrec', 'mse_log', 'slot')

    def __init__(self, z_in, z_out, H, W, sig_log=None, sigma_est=0.0, xrec=None, mse_log=None, slot=None):
        self.z_in, self.z_out, self.H, self.W = z_in, z_out, H, W
        self.sig_log, self.sigma_est, self.xrec, self.mse_log, self.slot = sig_log, sigma_est, xrec, mse_log, slot


def _require_device_objects(problem, denoiser):
    if not hasattr(problem, '_dev_grad'):
        raise TypeError('%s is not a pnp_svrg_b200 problem: the loops run on the GPU only (no CPU fallback)'
                        % type(problem).__name__)
    if not hasattr(denoiser, '_dev_denoise'):
        raise TypeError('%s is not a pnp_svrg_b200 denoiser: the loops run on the GPU only (no CPU fallback)'
                        % type(denoiser).__name__)


class HostDrawRing:
    """Look-ahead queue of host minibatch draws (mb_source='host') that land straight in a ring of staging buffers:
    a handle on the native queue of include/pnp_b200.h (``pnp_host_draws_*``).

    Draw number c is produced by one single-threaded run of the C sampler on a native worker thread and written
    into ``buffers[c % len(buffers)]`` -- the pinned buffer its host->device copy reads, so the loop never copies
    the indices and never touches Python threads.  ``ahead`` draws are kept in flight.  The sequence is the same as
    without threads: draw c is ``feistel_sample(n, B, seed, c)`` gathered through ``support``.
      next()   slot of the next draw (CPU only: the caller must be done with a buffer len(buffers) - ahead calls later)
      stage()  next + extras behind the indices + asynchronous copy to the device buffer + the event that guards the
               reuse of the staging buffer, in one native call"""

    def __init__(self, lib, n, B, seed, support, buffers, ahead, support_dev_ptr=None):
        import ctypes as C
        if support is not None and support_dev_ptr:
            raise ValueError('HostDrawRing: give the support list on the host or on the device, not both')
        for b in buffers:
            if b.dtype != np.int32 or b.ndim != 1 or b.size < B or not b.flags['C_CONTIGUOUS'] or not b.flags['WRITEABLE']:
                raise ValueError('HostDrawRing buffers must be writable contiguous int32 vectors of at least B entries')
        if support is not None and (support.dtype != np.int32 or not support.flags['C_CONTIGUOUS']):
            raise ValueError('HostDrawRing support must be a contiguous int32 vector')
        self.lib, self.B = lib, int(B)
        self.support, self.buffers = support, list(buffers)      # kept alive for the worker threads
        self.n_extra = min(b.size for b in self.buffers) - self.B
        self.drawn = 0                                             # draws handed out so far = number of the next one
        self._slot = C.c_int(0)
        self._extras = (C.c_int * max(self.n_extra, 1))()
        self._h = C.c_void_p()
        ptrs = (C.c_void_p * len(self.buffers))(*[b.ctypes.data for b in self.buffers])
        _lib.check(lib.pnp_host_draws_create(C.byref(self._h), int(n), self.B, int(seed) & 0xffffffff,
                                             None if support is None else support.ctypes.data, ptrs, len(self.buffers),
                                             int(ahead)))
        if support_dev_ptr:
            # the staging buffers then hold RANKS; the device maps them through its own copy of the list after each copy
            _lib.check(lib.pnp_host_draws_set_device_support(self._h, support_dev_ptr))

    def next(self):
        rc = self.lib.pnp_host_draws_next(self._h, self._slot)
        if rc:
            _lib.check(rc)
        self.drawn += 1
        return self._slot.value

    def stage(self, dst_dev_ptr, extras, stream_ptr):
        n = len(extras)
        if n > self.n_extra:
            raise ValueError('%d extra entries do not fit the staging buffers (%d)' % (n, self.n_extra))
        for i in range(n):
            self._extras[i] = int(extras[i])
        rc = self.lib.pnp_host_draws_stage(self._h, dst_dev_ptr, self._extras, n, stream_ptr, self._slot)
        if rc:
            _lib.check(rc)
        self.drawn += 1
        return self._slot.value

    def stage_many(self, dst_dev_ptr, n_draws, dst_stride, stream_ptr):
        """the next ``n_draws`` draws to ``dst[j * dst_stride + ...]`` in one native call (consecutive ring rows: one copy)"""
        rc = self.lib.pnp_host_draws_stage_many(self._h, dst_dev_ptr, int(n_draws), int(dst_stride), stream_ptr)
        if rc:
            _lib.check(rc)
        self.drawn += int(n_draws)

    def close(self):
        """Let the draws in flight finish (they write into the buffers) and stop the worker threads."""
        if self._h:
            self.lib.pnp_host_draws_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, problem, denoiser, mini_batch_size=0, mb_source='legacy', mb_seed=0, mb_stream=None,
                 fast=False, n_extra_ints=0):
        _require_device_objects(problem, denoiser)
        self.p, self.d = problem, denoiser
        self.lib = _lib.load()
        self.dev = problem._device
        self.H, self.W, self.N = problem.H, problem.W, problem.N
        self.B = int(mini_batch_size)
        self.mb_source = mb_source
        self.mb_stream = mb_stream
        self.fast = fast
        if mb_source not in ('legacy', 'host', 'device', 'stream'):
            raise ValueError("mb_source must be 'legacy', 'host', 'device' or 'stream'")
        if mb_source == 'stream' and mb_stream is None:
            raise ValueError("mb_source='stream' needs mb_stream (a sequence of index arrays)")
        n_meas = getattr(problem, 'M0', None) if getattr(problem, 'pname', '') == 'csmri' else getattr(problem, 'M', None)
        if self.B > 0 and mb_source in ('device', 'host') and n_meas is not None and self.B > int(n_meas):
            # problems/problem.py:112-115: the reference warns and np.random.choice(replace=False) raises
            raise ValueError('Cannot take a larger sample (%d) than the %d measurements' % (self.B, int(n_meas)))
        self.rng = np.random.default_rng(mb_seed) if mb_source == 'host' else None
        self.mb_seed = int(mb_seed)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.synchronize(self.dev)
        self.sptr = self.stream.cuda_stream
        with torch.cuda.stream(self.stream):
            self.z = D.to_lines(problem.Xinit, self.H, self.W, self.dev)
            self.mse_log = torch.zeros(LOG_CHUNK, dtype=torch.float64, device=self.dev)
            self.sig_log = torch.zeros(LOG_CHUNK, dtype=torch.float64, device=self.dev)
            self.counters = torch.zeros(4, dtype=torch.int32, device=self.dev)   # [0] log slot [1] mb cursor [2] draws
            self.step = torch.zeros(1, dtype=torch.float32, device=self.dev)
            self.sel = problem._dev_new_sel(self.B) if self.B > 0 else None
            self.n_extra = n_extra_ints
            if self.B > 0:
                if mb_source != 'host':          # 'host' stages through the ring of the look-ahead draws (below)
                    self.idx_host = torch.empty(self.B + n_extra_ints, dtype=torch.int32).pin_memory()
                self.idx_dev = torch.zeros(self.B + n_extra_ints, dtype=torch.int32, device=self.dev)
            self.barrier_ws = torch.zeros(2, dtype=torch.int32, device=self.dev)   # software grid barrier of chained launches
        self.chain = False          # set while an epoch graph is captured: passes use programmatic dependent launch
        self.sw_barrier = False     # set while an epoch graph is captured: the tail kernel uses the software grid barrier
        self.slot_ptr = self.counters[0:1]
        self.cursor_ptr = self.counters[1:2]
        self.draw_ptr = self.counters[2:3]
        self.slot_host = 0          # next log slot the device will write
        self.slot_base = 0          # first slot not yet read back
        self.n_prox = 0
        self.psnr_log = []
        self.sig_hist = []
        self.time_log = []
        self.gradient_time = 0.0
        self.denoise_time = 0.0
        self._stream_pos = 0
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        self._draws = None
        self._host_ring = None
        if mb_source == 'host' and self.B > 0:
            # native look-ahead queue: one single-threaded run of the C sampler per draw, several draws in flight
            # (a 100k-index draw takes 0.3-0.6 ms on one core), each written into the pinned buffer its host->device
            # copy reads; draw_host() stages the next one with a single native call.  One pinned allocation, sliced.
            # worker threads = draws in flight: a draw is ~0.1 ms of one core (ranks only), four workers make 40k draws/s
            # against the ~16k/s a B200 consumes at 2048^2; eight measured SLOWER through the public call (0.86-0.88 x the
            # device-resident rate vs 0.95-0.98 x with four: profiles/r02_e2e_workers.txt)
            ahead = int(os.environ.get('PNP_HOST_AHEAD', '0')) or max(2, min(4, ncpu // 2))
            # how far the host may run ahead.  A staging buffer is reused only after the copy that read it, and the copies of
            # an epoch wait on the device for the epoch two before it: with 12 extra buffers (two epochs of T2 = 10 draws) the
            # host staged every epoch late and the epoch graph waited 64 us for its index sets (66.4 vs 61.7 us per inner
            # iteration at 2048^2, scripts/prof_epoch_host.py); 32 keeps four epochs in hand
            depth = ahead + max(2, int(os.environ.get('PNP_HOST_RING_EXTRA', '32')))
            # (page-locked allocations cost milliseconds: the ring is kept on the problem object between calls; the
            # previous owner's queue has been closed by its result(), nothing writes into it any more)
            ring = D.pinned_buffer(('draw_ring', id(problem)), (depth, self.B + n_extra_ints), torch.int32)
            self._host_ring = [ring[i] for i in range(depth)]
            self.idx_host = self._host_ring[0]
            sup = getattr(problem, '_support_host', None)
            self._host_views = [t.numpy() for t in self._host_ring]
            sup_dev = getattr(problem, '_support', None)
            if sup is not None and sup_dev is not None and os.environ.get('PNP_HOST_GATHER', '0') != '1':
                # the host draws WHICH of the M0 measurements form the minibatch (distinct ranks); the device resolves rank ->
                # k-space position through the support list it holds anyway (the device sampler's), right behind the H2D
                # copy: the gather through a 5 MB table was 70 % of a 100k draw on one host core (cache misses)
                self._draws = HostDrawRing(self.lib, sup.size, self.B, self.mb_seed, None, self._host_views, ahead,
                                           support_dev_ptr=D.ptr(sup_dev))
                self._host_support = sup
            else:
                self._host_support = None
                self._draws = HostDrawRing(self.lib, problem.M if sup is None else sup.size, self.B, self.mb_seed, sup,
                                           self._host_views, ahead)
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self._pending = []          # fast mode: (kind,) markers for slots not yet read back
        self.graph = None
        self.fused_prox = None      # None: try the single-launch prox; False: image too large for it
        self.fused_tail = None      # same for the single-launch pass 3 + update + prox of CSMRI (algorithms/_loops.py)
        self.deferred = []
        self.since_sync = 0
        self.side = None
        self.sigma_ready = False    # set when the gradient pass already accumulated sigma for this slot
        self.uses_sigma = getattr(denoiser, '_uses_sigma_est', True)

    # ------------------------------------------------------------------ helpers
    def check(self, rc):
        if rc:
            _lib.check(rc)

    def set_step(self, value):
        self.step.fill_(float(value))

    def fork(self, side_fn, main_fn):
        """Run side_fn on a second stream concurrently with main_fn (both are captured as parallel
        graph branches when a capture is active), then join."""
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.dev)
            self._ev_fork = torch.cuda.Event()
            self._ev_join = torch.cuda.Event()
        self._ev_fork.record(self.stream)
        self.side.wait_event(self._ev_fork)
        with torch.cuda.stream(self.side):
            side_fn()
            self._ev_join.record(self.side)
        main_fn()
        self.stream.wait_event(self._ev_join)

    def copy(self, dst, src):
        self.check(self.lib.pnp_copy_f32(D.ptr(dst), D.ptr(src), dst.numel(), self.sptr))

    def psnr_of(self, z_lines):
        """Problem.PSNR of a device iterate (one reduction kernel + 8-byte readback)."""
        self.stream.synchronize()
        with torch.cuda.stream(self.stream):
            s = self.p._sq_err_dev(z_lines)
        return self.p._psnr_from_sum(s)

    def _psnr_from_sum(self, s):
        return self.p._psnr_from_sum(s)

    # ------------------------------------------------------------------ minibatches
    def draw_host(self, extra=()):
        """Draw one minibatch on the host in the reference's RNG call order and stage it in pinned
        memory; returns the positions (k-space / measurement indices)."""
        if self.mb_source == 'legacy':
            idx = self.p._draw_indices(self.B)
        elif self.mb_source == 'host':
            # the draws of the next iterations are produced by native worker threads while the GPU works on the
            # current one (same sequence as without the threads); the indices are already in the pinned buffer, and
            # the copy to idx_dev is enqueued here, on the engine stream, by the same native call
            slot = self._draws.stage(D.ptr(self.idx_dev), extra, self.sptr)
            self.idx_host = self._host_ring[slot]
            # (ranks into the support list when the device resolves them: self._host_support[...] gives the positions)
            return self._host_views[slot][:self.B]
        elif self.mb_source == 'stream':
            idx = np.asarray(self.mb_stream[self._stream_pos])
            self._stream_pos += 1
            if idx.size != self.B:
                raise ValueError('mb_stream entry %d has %d positions, expected %d' % (self._stream_pos - 1, idx.size, self.B))
        else:
            raise RuntimeError('draw_host called with mb_source=%r' % self.mb_source)
        buf = self.idx_host.numpy()
        buf[:self.B] = idx
        for i, e in enumerate(extra):
            buf[self.B + i] = e
        return idx

    def next_host_buffer(self):
        """Rotate to the next pinned staging buffer (``self.idx_host``), waiting only for the H2D copy that read it
        last time round.  (Not used by mb_source='host': its draws own their ring, see HostDrawRing.)"""
        if self._host_ring is None:
            self._host_ring = [self.idx_host] + [torch.empty_like(self.idx_host).pin_memory() for _ in range(3)]
            self._host_ev = [None] * len(self._host_ring)
            self._host_pos = 0
        self._host_pos = (self._host_pos + 1) % len(self._host_ring)
        ev = self._host_ev[self._host_pos]
        if ev is not None:
            ev.synchronize()
        self.idx_host = self._host_ring[self._host_pos]

    def mark_host_buffer(self):
        """Record that the copy just enqueued on the engine stream reads the current staging buffer."""
        ev = self._host_ev[self._host_pos] or torch.cuda.Event()
        ev.record(self.stream)
        self._host_ev[self._host_pos] = ev

    def upload_sel(self):
        """pinned -> device copy of the staged minibatch and rebuild of the selection."""
        if self._draws is None:                  # mb_source='host': draw_host() has already enqueued the copy
            self.idx_dev.copy_(self.idx_host, non_blocking=True)
        # the selection buffer is zero on entry: the gradient pass that consumes it clears it again
        self.p._dev_set_sel(self.sel, self.idx_dev, self.B, clear=False)

    def sample_sel_device(self):
        self.p._dev_sample_sel(self.sel, self.B, self.mb_seed, counter=self.draw_ptr, clear=False)

    def sel_job(self):
        """Descriptor of the current minibatch for problems whose gradient pass builds the selection itself
        (CSMRI: ``_dev_grad(..., sel_job=)``, no separate selection launch); None otherwise."""
        if self.B <= 0 or not getattr(self.p, '_inpass_sel', False):
            return None
        if self.mb_source == 'device':
            return dict(count=self.B, idx=None, seed=self.mb_seed, counter=self.draw_ptr)
        return dict(count=self.B, idx=self.idx_dev, cursor=None)

    # ------------------------------------------------------------------ prox + log
    def prox(self, z_in, z_out):
        """sigma estimate + denoiser + squared error against the ground truth into the current slot."""
        if self.uses_sigma and not self.sigma_ready and self.fused_prox is not False and hasattr(self.d, '_dev_prox_fused'):
            ok = self.d._dev_prox_fused(ProxCtx(z_in, z_out, self.H, self.W, sig_log=self.sig_log, xrec=self.p._xrec_dev,
                                                mse_log=self.mse_log, slot=self.slot_ptr))
            self.fused_prox = ok
            if ok:
                return
        if self.uses_sigma and not self.sigma_ready:
            self.check(self.lib.pnp_estimate_sigma(D.ptr(z_in), self.H, self.W, 1, D.ptr(self.sig_log),
                                                   D.ptr(self.slot_ptr), self.sptr))
        self.sigma_ready = False
        self.d._dev_denoise(ProxCtx(z_in, z_out, self.H, self.W, sig_log=self.sig_log if self.uses_sigma else None,
                                    xrec=self.p._xrec_dev, mse_log=self.mse_log, slot=self.slot_ptr))
        if not getattr(self.d, '_fused_psnr', True):
            self.check(self.lib.pnp_sq_err(D.ptr(z_out), D.ptr(self.p._xrec_dev), self.N, 1, D.ptr(self.mse_log),
                                           D.ptr(self.slot_ptr), self.sptr))

    def advance(self, n=3):
        self.check(self.lib.pnp_advance(D.ptr(self.counters), n, self.sptr))

    def read_slot(self):
        """faithful mode: PSNR of the prox output just computed (synchronises the stream)."""
        s = self.slot_host
        vals = torch.stack([self.mse_log[s], self.sig_log[s]]).cpu().numpy()
        self.slot_host += 1
        self.slot_base = self.slot_host
        self.n_prox += 1
        if self.slot_host >= LOG_CHUNK:
            self._reset_logs()
        self.sig_hist.append(float(vals[1]) / self.W)
        return self._psnr_from_sum(float(vals[0]))

    def _reset_logs(self):
        self.mse_log.zero_()
        self.sig_log.zero_()
        self.counters[0:1].zero_()
        self.slot_host = 0
        self.slot_base = 0

    # deferred logging (fast mode without stop rules): iterations are only counted while they are
    # enqueued; PSNR values are read back in one go by resolve()
    def defer_slots(self, n):
        if not self.deferred:
            self._defer_t0 = time.time()
        self.deferred.append(n)
        self.slot_host += n
        self.since_sync += n

    def defer_dup(self):
        """the log repeats the previous value (PnP-SVRG logs PSNR(z) again at every snapshot)"""
        self.deferred.append('dup')

    def resolve(self):
        if not self.deferred:
            return
        vals = self.flush_fast()
        n_slots = sum(d for d in self.deferred if d != 'dup')
        dt = (time.time() - self._defer_t0) / max(n_slots, 1)
        it = iter(vals)
        for d in self.deferred:
            if d == 'dup':
                self.psnr_log.append(self.psnr_log[-1])
                self.time_log.append(0.0)
            else:
                for _ in range(d):
                    self.psnr_log.append(next(it))
                    self.time_log.append(dt)
                    self.n_prox += 1
        self.deferred = []
        self.since_sync = 0

    def flush_fast(self):
        """fast mode: read back every slot written since the last flush."""
        self.stream.synchronize()
        lo, n = self.slot_base, self.slot_host
        if n == lo:
            return []
        both = torch.stack([self.mse_log[lo:n], self.sig_log[lo:n]]).cpu().numpy()     # one read-back for both logs
        mse, sig = both[0], both[1]
        self.sig_hist.extend((sig / self.W).tolist())
        with torch.cuda.stream(self.stream):
            self._reset_logs()
        return list(np.atleast_1d(self._psnr_from_sum(mse)))          # vectorised: same arithmetic as per value

    # ------------------------------------------------------------------ graphs
    def capture(self, fn):
        """Capture the launches issued by fn() on the engine stream into an executable graph."""
        import ctypes as C
        self.stream.synchronize()
        self.check(self.lib.pnp_graph_begin(self.sptr))
        try:
            with torch.cuda.stream(self.stream):
                fn()
        finally:
            exec_ = C.c_void_p()
            rc = self.lib.pnp_graph_end(self.sptr, C.byref(exec_))
        self.check(rc)
        return exec_

    def replay(self, exec_):
        self.check(self.lib.pnp_graph_launch(exec_, self.sptr))

    def destroy(self, exec_):
        if exec_:
            self.lib.pnp_graph_destroy(exec_)

    def result(self, name):
        if self._draws is not None:
            self._draws.close()
        self.resolve()
        self.stream.synchronize()
        z = D.from_lines(self.z, self.H, self.W)
        return {
            'z': z,
            'time_per_iter': self.time_log,
            'psnr_per_iter': self.psnr_log,
            'gradient_time': self.gradient_time,
            'denoise_time': self.denoise_time,
            'algo_name': name,
        }


class Budget:
    """Loop bound: the reference's wall clock ``tt`` plus the additive ``max_iters`` (prox calls)."""

    def __init__(self, tt, max_iters):
        self.tt = float(tt)
        self.max_iters = None if max_iters is None else int(max_iters)
        self.t0 = time.time()
        self.calls = 0

    def alive(self):
        if self.max_iters is not None and self.calls >= self.max_iters:
            return False
        return (time.time() - self.t0) < self.tt

    def left(self):
        return None if self.max_iters is None else self.max_iters - self.calls


def stop_rule(start_psnr, last_psnr, converge_check, diverge_check):
    if converge_check is True and np.abs(start_psnr - last_psnr) < TOL:
        return True
    if diverge_check is True and last_psnr < 0:
        return True
    return False


# ---------------------------------------------------------------------------------------------
# Host twin of the device minibatch sampler (csrc/csmri.cuh::feistel_perm): the same keyed
# cycle-walking Feistel permutation, vectorised in NumPy.  O(B) instead of the O(M0) permutation
# behind np.random.choice(..., replace=False); used by mb_source='host'.
def _mix32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def feistel_key(seed, counter, img=0):
    with np.errstate(over='ignore'):
        c = _mix32(np.array([(np.uint64(counter) * np.uint64(0x632be5ab) + np.uint64(img)) & np.uint64(0xffffffff)],
                            dtype=np.uint64).astype(np.uint32))
        return int(_mix32(np.uint32(seed & 0xffffffff) ^ c)[0])


def feistel_sample(n, count, seed, counter, img=0):
    """positions feistel_perm(i, n, key) for i in [0, count) -- `count` distinct values in [0, n)."""
    hb = 1
    while (1 << (2 * hb)) < n:
        hb += 1
    hm = np.uint32((1 << hb) - 1)
    b = np.uint32((n + (1 << hb) - 1) >> hb)         # unbalanced domain 2^hb x ceil(n / 2^hb), see csrc/csmri.cuh
    key = np.uint32(feistel_key(seed, counter, img))
    x = np.arange(count, dtype=np.uint32)
    out = np.empty(count, dtype=np.uint32)
    pending = np.arange(count)
    rk = lambda rd: np.uint32((int(key) + 0x9e3779b9 * rd) & 0xffffffff)
    with np.errstate(over='ignore'):
        while pending.size:
            l, r = x >> np.uint32(hb), x & hm
            for rd in (0, 2):
                f0 = ((_mix32(r ^ rk(rd + 1)).astype(np.uint64) * np.uint64(b)) >> np.uint64(32)).astype(np.uint32)
                t = l + f0
                t = np.where(t >= b, t - b, t)
                l, r = r, t
                f1 = _mix32(r ^ rk(rd + 2)) & hm
                l, r = r, (l + f1) & hm
            x = (l << np.uint32(hb)) | r
            ok = x < np.uint32(n)
            out[pending[ok]] = x[ok]
            pending, x = pending[~ok], x[~ok]
    return out.astype(np.int64)
