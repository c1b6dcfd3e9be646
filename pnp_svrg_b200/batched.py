"""Batched PnP-SVRG for sweeps: many same-size CSMRI reconstructions advance together, one kernel
launch per pass for the whole batch (the kernels take a batch dimension, blockIdx.y = problem).

A 256x256 iteration moves ~2 MB and is launch-latency bound on a B200; the sweeps of
script_diff_sampratio_set12.py / script_diff_snr_set12.py (image x sampling ratio x SNR, independent
reconstructions, `Pool.map` in the reference) are therefore run as batches: per-problem masks,
measurements, ground truths, M0 and step sizes live in stacked device arrays, the minibatch of every
problem is drawn by the device sampler (keyed by problem index), and the PSNR / sigma logs are
[slot][problem] arrays read back once at the end.

Only the paper-mode SVRG + wavelet-prox combination (config 1 / 4) is batched here; every other
combination goes through the per-problem engine.
"""
import ctypes as C
import os
import time

import numpy as np
import torch

from . import _lib, device as D


def csmri_host_spec(image, H, W, sample_prob, snr, rng=np.random):
    """Host-side construction of one CSMRI problem, same draws and arithmetic as problems.CSMRI
    (reference problems/CSMRI.py:12-41) but without touching the GPU.  `rng` is np.random or a
    np.random.RandomState (thread-safe sweeps)."""
    from .problems.problem import load_image
    xrec = load_image(None, image, H, W)
    mask = rng.choice([0, 1], size=(H, W), p=[1 - sample_prob, sample_prob])
    Y0 = mask * np.fft.fft2(xrec)
    sigma = np.sqrt(np.linalg.norm(Y0.ravel()) / 10 ** (snr / 10) / H / W)
    Y = Y0 + mask * rng.normal(0, sigma, Y0.shape)
    x0 = np.absolute(np.fft.ifft2(Y))
    xinit = (x0 - x0.min()) / (x0.max() - x0.min())
    hp = H // 2
    Ym = mask * Y
    Ymir = np.conj(Ym[(-np.arange(H)) % H][:, (-np.arange(W)) % W])
    return dict(H=H, W=W, M0=int(np.count_nonzero(mask)), sigma=float(sigma),
                xrec=np.ascontiguousarray(xrec.T, dtype=np.float32), xinit=np.ascontiguousarray(xinit.T, dtype=np.float32),
                Y1=Ym[:hp].astype(np.complex64), Y2=Ymir[:hp].astype(np.complex64),
                Y1n=Ym[hp].astype(np.complex64), Y2n=Ymir[hp].astype(np.complex64),
                support=np.flatnonzero(mask).astype(np.int32), data_range=1.0 if xrec.min() >= 0 else 2.0)


_IMAGE_CACHE = {}


_DEV_IMAGE_CACHE = {}


def _device_lines(im, H, W, dev):
    """normalised ground truth of one source image as a device tensor in line layout [W][H]; sweeps revisit the same few
    images hundreds of times, so each one is normalised and uploaded once per process and device"""
    from .problems.problem import load_image
    key = (id(im), H, W, str(dev))
    hit = _DEV_IMAGE_CACHE.get(key)
    if hit is None or hit[0] is not im:
        x = load_image(None, im, H, W).astype(np.float32)
        hit = _DEV_IMAGE_CACHE[key] = (im, torch.from_numpy(np.ascontiguousarray(x.T)).to(dev), bool(x.min() >= 0))
    return hit


def csmri_device_batch_native(images, sample_probs, snrs, H, W, seed=0, device=None, sync=True):
    """Construction of a whole batch of CSMRI problems on the device with the package's own kernels
    (pnp_csmri_build_batch, csrc/build.cuh): same model as problems/CSMRI.py:12-41 + problems/problem.py:58-61 --
    Bernoulli(p) mask, Y = mask o (fft2(X) + N(0, sigma)), sigma from the SNR, Xinit = minmax(|ifft2(Y)|), ascending
    support lists -- with counter-based random numbers and the iteration's own FFT passes; no torch.fft / torch.sort /
    torch RNG.  ``sync=False``: nothing is read back (``m0_host`` is None; M0, 1 / M0 and sigma stay device tensors), so a
    sweep can build batch k + 1 without ever waiting for the device."""
    import ctypes as C
    dev = device or D.require_cuda()
    lib = _lib.load()
    nb, hp, N = len(images), H // 2, H * W
    hits = [_device_lines(im, H, W, dev) for im in images]
    xrec = torch.stack([h[1] for h in hits])                                     # [nb][W][H]
    p = torch.from_numpy(np.asarray(sample_probs, dtype=np.float32)).to(dev, non_blocking=True)
    snr = torch.from_numpy(np.asarray(snrs, dtype=np.float32)).to(dev, non_blocking=True)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = dict(device=True, native=True, nb=nb, H=H, W=W, xrec=xrec, xinit=f32(nb, W, H),
               Y1=f32(nb, hp, W, 2), Y2=f32(nb, hp, W, 2), Y1n=f32(nb, W, 2), Y2n=f32(nb, W, 2),
               bits_full=torch.empty((nb, hp, W), dtype=torch.uint8, device=dev),
               m0=torch.empty(nb, dtype=torch.int32, device=dev), inv_m0=f32(nb), sigma_dev=f32(nb),
               # the lists are N wide (M0 is only known on the device); entries past M0 are never read
               support=torch.empty((nb, N), dtype=torch.int32, device=dev),
               data_range=np.array([1.0 if h[2] else 2.0 for h in hits]))
    work = torch.empty(int(lib.pnp_csmri_build_batch_workspace(H, W, nb)), dtype=torch.uint8, device=dev)
    args = _lib.CsmriBuildArgs(
        H=H, W=W, batch=nb, seed=int(seed) & 0xffffffff, x=D.ptr(xrec), p=D.ptr(p), snr=D.ptr(snr),
        bits_full=D.ptr(out['bits_full']), m0=D.ptr(out['m0']), inv_m0=D.ptr(out['inv_m0']), support=D.ptr(out['support']),
        support_img_stride=N, Y1=D.ptr(out['Y1']), Y2=D.ptr(out['Y2']), Y1n=D.ptr(out['Y1n']), Y2n=D.ptr(out['Y2n']),
        xinit=D.ptr(out['xinit']), sigma=D.ptr(out['sigma_dev']), work=D.ptr(work))
    _lib.check(lib.pnp_csmri_build_batch(C.byref(args), D.stream()))
    out['_keep'] = (p, snr, work)                    # alive until the stream has run the launches above
    if sync:
        out['m0_host'] = out['m0'].cpu().numpy().astype(np.int32)
        out['sigma'] = out['sigma_dev'].cpu().numpy()
    else:
        out['m0_host'] = None
        out['sigma'] = None
    return out


def csmri_device_batch(images, sample_probs, snrs, H, W, seed=0, device=None, native=None, sync=True):
    """A batch of CSMRI problems built on the device: the package's own constructor (csmri_device_batch_native) unless
    ``native=False`` / PNP_BUILD_TORCH=1 asks for the round-1 torch.fft one below (kept as a cross-check)."""
    if native is None:
        native = os.environ.get('PNP_BUILD_TORCH', '0') != '1'
    if native:
        return csmri_device_batch_native(images, sample_probs, snrs, H, W, seed=seed, device=device, sync=sync)
    return csmri_device_batch_torch(images, sample_probs, snrs, H, W, seed=seed, device=device)


def csmri_device_batch_torch(images, sample_probs, snrs, H, W, seed=0, device=None):
    """Construction of a whole batch of CSMRI problems ON THE DEVICE (the step before the hot path, SURVEY section
    8(f) rank 2): same model as problems/CSMRI.py:12-41 -- Bernoulli(p) mask, Y = mask o (fft2(X) + N(0, sigma)) with
    sigma from the SNR, Xinit = minmax(|ifft2(Y)|) -- but drawn with a torch CUDA generator and transformed with
    torch.fft in float32, so the draws differ from the NumPy ones of `csmri_host_spec` (sweeps are unseeded in the
    reference).  Library calls are confined to this constructor; the iterations run on the package's own kernels.
    Returns the stacked device tensors BatchedSVRG takes in place of a list of host specs."""
    from .problems.problem import load_image
    dev = device or D.require_cuda()
    nb, hp = len(images), H // 2
    def normalised(im):                       # sweeps revisit the same few images: normalise each array once
        key = (id(im), H, W)
        hit = _IMAGE_CACHE.get(key)
        if hit is None or hit[0] is not im:
            hit = _IMAGE_CACHE[key] = (im, load_image(None, im, H, W).astype(np.float32))
        return hit[1]
    x = torch.from_numpy(np.stack([normalised(im) for im in images])).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    p = torch.tensor(np.asarray(sample_probs, dtype=np.float32), device=dev).view(nb, 1, 1)
    snr = torch.tensor(np.asarray(snrs, dtype=np.float32), device=dev)
    mask = torch.rand((nb, H, W), generator=gen, device=dev) < p
    y0 = torch.fft.fft2(x) * mask
    # problems/problem.py:58-61: sigma = sqrt(||Y0||_2 / 10^(snr/10) / H / W)   (norm, not norm^2, as in the reference)
    sigma = torch.sqrt(torch.linalg.vector_norm(y0.reshape(nb, -1), dim=1) / 10 ** (snr / 10) / H / W)
    y = y0 + mask * (torch.randn((nb, H, W), generator=gen, device=dev) * sigma.view(nb, 1, 1))      # real noise, CSMRI.py:32-33
    x0 = torch.fft.ifft2(y).abs()
    lo, hi = x0.amin(dim=(1, 2), keepdim=True), x0.amax(dim=(1, 2), keepdim=True)
    xinit = (x0 - lo) / (hi - lo)
    ymir = torch.conj(torch.roll(torch.flip(y, dims=(1, 2)), shifts=(1, 1), dims=(1, 2)))              # conj Y[-ky][-kx]
    c = lambda t: torch.view_as_real(t.to(torch.complex64).contiguous()).contiguous()
    m0 = mask.reshape(nb, -1).sum(dim=1).to(torch.int32)
    # padded support lists: indices of the sampled positions first (ascending), as np.flatnonzero gives them
    N = H * W
    key = torch.where(mask.reshape(nb, N), torch.arange(N, device=dev, dtype=torch.int32).expand(nb, N),
                      torch.full((1,), N, device=dev, dtype=torch.int32).expand(nb, N))
    m0_host = m0.cpu().numpy()
    stride = int(m0_host.max())
    support = torch.sort(key, dim=1).values[:, :stride].contiguous()
    support = torch.where(support < N, support, torch.zeros_like(support))
    # selection bytes of the full mask in the kernels' packed layout [H/2][W] (csrc/csmri.cuh::set_sel_bits):
    # bit0 sel[ky][kx], bit1 sel[-ky][-kx]; row 0 also carries the Nyquist row in bits 2 and 3
    m8 = mask.to(torch.uint8)
    mir = torch.roll(torch.flip(m8, dims=(1, 2)), shifts=(1, 1), dims=(1, 2))
    bits = m8[:, :hp] + 2 * mir[:, :hp]
    bits[:, 0] += 4 * m8[:, hp] + 8 * mir[:, hp]
    return dict(device=True, nb=nb, H=H, W=W, m0_host=m0_host.astype(np.int32), m0=m0.contiguous(), support=support,
                bits_full=bits.contiguous(),
                xrec=x.transpose(1, 2).contiguous(), xinit=xinit.transpose(1, 2).contiguous(),
                Y1=c(y[:, :hp]), Y2=c(ymir[:, :hp]), Y1n=c(y[:, hp]), Y2n=c(ymir[:, hp]),
                sigma=sigma.cpu().numpy(), data_range=np.ones(nb))


class BatchedSVRG:
    """PnP-SVRG (paper-mode VR, algorithms/pnp_svrg.py:8-105 with line 53) + wavelet prox
    (denoisers/TV.py) on a batch of CSMRI problems of one size."""

    def __init__(self, specs, T2, mini_batch_size, etas, seed=0, lr_decay=1.0, sigma_modifier=1.0, max_slots=4096):
        self.lib = _lib.load()
        self.dev = D.require_cuda()
        on_dev = isinstance(specs, dict) and specs.get('device')          # csmri_device_batch() output
        if on_dev:
            self.nb = nb = specs['nb']
            self.H, self.W = specs['H'], specs['W']
            m0_all = specs['m0_host']                  # None: built without a read-back (csmri_device_batch(sync=False))
        else:
            self.nb = nb = len(specs)
            self.H, self.W = specs[0]['H'], specs[0]['W']
            if any(s['H'] != self.H or s['W'] != self.W for s in specs):
                raise ValueError('all problems of a batch must have the same size')
            m0_all = np.array([s['M0'] for s in specs], dtype=np.int32)
        self.N = N = self.H * self.W
        hp = self.H // 2
        self.T2, self.B, self.seed, self.lr_decay = int(T2), int(mini_batch_size), int(seed), float(lr_decay)
        self.sigma_modifier = float(sigma_modifier)
        if m0_all is not None and (m0_all < self.B).any():
            raise ValueError('mini_batch_size exceeds the number of measurements of a problem')
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sptr = self.stream.cuda_stream
        dev = self.dev

        def stack(key, dtype):
            return torch.from_numpy(np.ascontiguousarray(np.stack([s[key] for s in specs]))).to(dev).to(dtype).contiguous()

        def stack_c(key):
            return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(np.stack([s[key] for s in specs]))).to(dev)).contiguous()
        if on_dev:
            torch.cuda.current_stream(dev).synchronize()                 # the batch was built on the current stream
        with torch.cuda.stream(self.stream):
            self.m0_host = m0_all
            self.sup_stride = N if on_dev else int(self.m0_host.max())      # device batches: fixed stride, the object is reusable
            if on_dev:
                for v in specs.values():
                    if isinstance(v, torch.Tensor) and v.is_cuda:
                        v.record_stream(self.stream)
                self.xrec, self.z = specs['xrec'].clone(), specs['xinit'].clone()
                self.Y1, self.Y2, self.Y1n, self.Y2n = (specs[k].clone() for k in ('Y1', 'Y2', 'Y1n', 'Y2n'))
                self.m0 = specs['m0'].clone()
                self.support = torch.zeros((nb, N), dtype=torch.int32, device=dev)
                self.support[:, :specs['support'].shape[1]] = specs['support']
            else:
                self.xrec = stack('xrec', torch.float32)
                self.z = stack('xinit', torch.float32)
                self.Y1, self.Y2, self.Y1n, self.Y2n = stack_c('Y1'), stack_c('Y2'), stack_c('Y1n'), stack_c('Y2n')
                self.m0 = torch.from_numpy(self.m0_host).to(dev)
                sup = np.zeros((nb, self.sup_stride), dtype=np.int32)
                for i, s in enumerate(specs):
                    sup[i, :s['M0']] = s['support']
                self.support = torch.from_numpy(sup).to(dev)
            self.w = torch.empty_like(self.z)
            self.mu = torch.empty_like(self.z)
            self.S = torch.empty(nb * N, dtype=torch.float32, device=dev)
            self.bits_full = torch.zeros(nb * self.W * hp, dtype=torch.uint8, device=dev)
            self.bits_mb = torch.zeros(nb * self.W * hp, dtype=torch.uint8, device=dev)
            if on_dev and specs.get('inv_m0') is not None:
                self.inv_m0 = specs['inv_m0'].clone()
            else:
                self.inv_m0 = torch.from_numpy((1.0 / self.m0_host).astype(np.float32)).to(dev)
            self.step = torch.empty(nb, dtype=torch.float32, device=dev)
            self._set_etas(etas)
            self.mse_log = torch.zeros(max_slots * nb, dtype=torch.float64, device=dev)
            self.sig_log = torch.zeros(max_slots * nb, dtype=torch.float64, device=dev)
            self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
            if on_dev:
                self.bits_full.copy_(specs['bits_full'].reshape(-1))
            else:
                self.check(self.lib.pnp_csmri_sel_from_indices(D.ptr(self.bits_full), self.H, self.W, nb, D.ptr(self.support), 0,
                                                               self.sup_stride, None, 1, self.sptr))
                # full-mask bits: every problem's own support (counts differ, so one launch per distinct problem)
                for i in range(nb):
                    self.check(self.lib.pnp_csmri_sel_from_indices(
                        self.bits_full.data_ptr() + i * self.W * hp, self.H, self.W, 1,
                        self.support.data_ptr() + 4 * i * self.sup_stride, int(self.m0_host[i]), 0, None, 0, self.sptr))
            self.mse0 = torch.zeros(nb, dtype=torch.float64, device=dev)
            self.check(self.lib.pnp_sq_err(D.ptr(self.z), D.ptr(self.xrec), N, nb, D.ptr(self.mse0), None, self.sptr))
        self.data_range = np.asarray(specs['data_range'], dtype=np.float64) if on_dev else np.array([s['data_range'] for s in specs])
        self.max_slots = max_slots
        self.slots_used = 0
        self.graph = None
        self.side = torch.cuda.Stream(device=self.dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        self.outer = 0
        self.fused_prox = True
        self.whole_run_graph = False     # sweeps set it: capture all iterations of a run as one graph (see _capture_run)
        # the first inner iteration of an epoch skips its transform passes (whole-run graphs; PNP_BATCH_SKIP_FIRST=0: off, for A/B)
        self.skip_first = os.environ.get('PNP_BATCH_SKIP_FIRST', '1') != '0'
        # 128^2 / 256^2, and no more problems than clusters fit the device (15 on a B200): the whole run is ONE launch, one
        # thread-block cluster per problem (csrc/small.cuh).  Larger batches would run in waves of 15 clusters on 120 of
        # the 148 SMs; the three-pass kernels with the batch dimension (and two batches in flight, sweep.DeviceBatchPipeline)
        # then use the device better.  PNP_SMALL=0: never, PNP_SMALL=2: always (tests, measurements).
        mode = os.environ.get('PNP_SMALL', '1')
        self.use_small = (mode != '0' and self.lib.pnp_csmri_svrg_small_supported(int(self.H), int(self.W)) == 1
                          and (mode == '2' or nb <= self.lib.pnp_csmri_svrg_small_capacity(int(self.H), int(self.W))))
        self._run_graph, self._run_graph_n = None, None

    def check(self, rc):
        if rc:
            _lib.check(rc)

    def _set_etas(self, etas):
        """step sizes per problem: host numbers, or a device tensor (sweeps derive them from M0 without a read-back)"""
        if isinstance(etas, torch.Tensor):
            self.eta_host, self.eta_dev = None, etas.to(torch.float32).reshape(self.nb)
            self.step.copy_(self.eta_dev)
        else:
            self.eta_host = np.broadcast_to(np.asarray(etas, dtype=np.float64), (self.nb,)).copy()
            self.eta_dev = None
            self.step.copy_(torch.from_numpy(self.eta_host.astype(np.float32)), non_blocking=False)

    def _decayed_step(self):
        if self.eta_host is None:
            self.step.copy_(self.eta_dev * float(self.lr_decay ** self.outer))
        else:
            self.step.copy_(torch.from_numpy((self.eta_host * self.lr_decay ** self.outer).astype(np.float32)), non_blocking=True)

    def reload(self, batch, etas):
        """Load another device-built batch of the same shape into the existing buffers: allocations and the
        captured iteration graph are reused (sweeps run hundreds of batches)."""
        if not (isinstance(batch, dict) and batch.get('device')) or batch['nb'] != self.nb or (batch['H'], batch['W']) != (self.H, self.W):
            raise ValueError('reload needs a csmri_device_batch of the same shape')
        if batch['m0_host'] is not None and (batch['m0_host'] < self.B).any():
            raise ValueError('mini_batch_size exceeds the number of measurements of a problem')
        # the batch was built on the current stream: this engine's stream waits for it on the device, the host does not
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            for v in batch.values():                       # read on this engine's stream: not to be recycled before that
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(self.stream)
            self.xrec.copy_(batch['xrec']); self.z.copy_(batch['xinit'])
            for k in ('Y1', 'Y2', 'Y1n', 'Y2n'):
                getattr(self, k).copy_(batch[k])
            self.m0.copy_(batch['m0'])
            self.m0_host = batch['m0_host']
            self.support[:, :batch['support'].shape[1]] = batch['support']
            self.bits_full.copy_(batch['bits_full'].reshape(-1))
            if batch.get('inv_m0') is not None:
                self.inv_m0.copy_(batch['inv_m0'])
            else:
                self.inv_m0.copy_(torch.from_numpy((1.0 / self.m0_host).astype(np.float32)), non_blocking=False)
            self._set_etas(etas)
            self.mse_log.zero_(); self.sig_log.zero_(); self.mse0.zero_()
            self.counters[0:2].zero_()                      # log slot and cursor; the draw counter keeps running
            self.check(self.lib.pnp_sq_err(D.ptr(self.z), D.ptr(self.xrec), self.N, self.nb, D.ptr(self.mse0), None, self.sptr))
        self.data_range = np.asarray(batch['data_range'], dtype=np.float64)
        self.slots_used = 0
        self.outer = 0

    def build_from_images(self, images, sample_probs, snrs, seed, eta_scale, eta_cap):
        """The next batch of a sweep built STRAIGHT INTO this engine's buffers by the package's own constructor
        (pnp_csmri_build_batch) on this engine's stream: no allocation, no copy, no read-back -- the host only enqueues, and
        the build overlaps whatever other engines are running.  Step sizes min(eta_scale * M0, eta_cap) are formed on the
        device from the M0 the constructor counted.  The engine must hold device-built problems (support lists N wide)."""
        import ctypes as C
        if len(images) != self.nb or self.sup_stride != self.N:
            raise ValueError('build_from_images needs %d images and an engine created from a device-built batch' % self.nb)
        hits = [_device_lines(im, self.H, self.W, self.dev) for im in images]
        self.prepare_build()
        if self._build_ps_ev is not None:
            self._build_ps_ev.synchronize()                 # the copy that last read the pinned parameters (long done)
        self._build_ps_host[0] = torch.from_numpy(np.asarray(sample_probs, dtype=np.float32))
        self._build_ps_host[1] = torch.from_numpy(np.asarray(snrs, dtype=np.float32))
        with torch.cuda.stream(self.stream):
            self._build_ps.copy_(self._build_ps_host, non_blocking=True)
            self._build_ps_ev = self._build_ps_ev or torch.cuda.Event()
            self._build_ps_ev.record(self.stream)
            torch.stack([h[1] for h in hits], out=self.xrec.view(self.nb, self.W, self.H))
            args = _lib.CsmriBuildArgs(
                H=self.H, W=self.W, batch=self.nb, seed=int(seed) & 0xffffffff, x=D.ptr(self.xrec), p=D.ptr(self._build_ps[0]),
                snr=D.ptr(self._build_ps[1]), bits_full=D.ptr(self.bits_full), m0=D.ptr(self.m0), inv_m0=D.ptr(self.inv_m0),
                support=D.ptr(self.support), support_img_stride=self.N, Y1=D.ptr(self.Y1), Y2=D.ptr(self.Y2), Y1n=D.ptr(self.Y1n),
                Y2n=D.ptr(self.Y2n), xinit=D.ptr(self.z), sigma=D.ptr(self.sigma_dev), work=D.ptr(self._build_work))
            self.check(self.lib.pnp_csmri_build_batch(C.byref(args), self.sptr))
            self.m0_host = None
            self._set_etas(torch.clamp(self.m0.to(torch.float32) * float(eta_scale), max=float(eta_cap)))
            self.mse_log.zero_(); self.sig_log.zero_(); self.mse0.zero_()
            self.counters[0:2].zero_()
            self.check(self.lib.pnp_sq_err(D.ptr(self.z), D.ptr(self.xrec), self.N, self.nb, D.ptr(self.mse0), None, self.sptr))
        self.data_range = np.array([1.0 if h[2] else 2.0 for h in hits])
        self.slots_used = 0
        self.outer = 0

    def prepare_build(self):
        """workspace, parameter buffers (one of them pinned host memory: milliseconds to allocate) of build_from_images;
        sweeps call it when they create an engine, so that no group of a sweep pays for it"""
        if getattr(self, '_build_work', None) is None:
            self._build_work = torch.empty(int(self.lib.pnp_csmri_build_batch_workspace(self.H, self.W, self.nb)), dtype=torch.uint8,
                                           device=self.dev)
            self._build_ps = torch.empty((2, self.nb), dtype=torch.float32, device=self.dev)
            self._build_ps_host = torch.empty((2, self.nb), dtype=torch.float32).pin_memory()
            self._build_ps_ev = None
            self.sigma_dev = torch.empty(self.nb, dtype=torch.float32, device=self.dev)

    # ---- the launches -------------------------------------------------------------------------
    def _grad(self, a, b, bits, with_y, stream, phases=0, **kw):
        args = _lib.CsmriGradArgs(
            H=self.H, W=self.W, batch=self.nb, a=D.ptr(a), b=D.ptr(b), S=D.ptr(self.S) if kw.get('spectrum', True) else None,
            bits=D.ptr(bits),
            Y1=D.ptr(self.Y1) if with_y else None, Y2=D.ptr(self.Y2) if with_y else None,
            Y1n=D.ptr(self.Y1n) if with_y else None, Y2n=D.ptr(self.Y2n) if with_y else None,
            gscale=float(kw.get('gscale', 1.0)), gscale_ptr=D.ptr(kw.get('gscale_ptr')), step=0.0,
            step_ptr=D.ptr(kw.get('step_ptr')), g_out=D.ptr(kw.get('g_out')), vadd=D.ptr(kw.get('vadd')), v_out=None,
            z_in=D.ptr(kw.get('z_in')), z_out=D.ptr(kw.get('z_out')), phases=int(phases),
            clear_bits=int(kw.get('clear', False)))
        self.check(self.lib.pnp_csmri_grad(C.byref(args), stream))

    def _snapshot(self):
        # mu = grad_full(z) (per-problem 1/M0), w = z                     (pnp_svrg.py:32-35)
        self._grad(self.z, None, self.bits_full, True, self.sptr, gscale_ptr=self.inv_m0, g_out=self.mu)
        self.check(self.lib.pnp_copy_f32(D.ptr(self.w), D.ptr(self.z), self.nb * self.N, self.sptr))

    def _inner(self, first_of_epoch=False):
        slot, draws = self.counters[0:1], self.counters[2:3]
        gk = dict(gscale=1.0 / self.B, vadd=self.mu, step_ptr=self.step, z_in=self.z, z_out=self.z, clear=True)
        if first_of_epoch and self.skip_first:
            # z == w bit for bit (the snapshot has just copied it): g_B(z) - g_B(w) is exactly zero, v = mu -- the selection,
            # the forward line pass and the column pass are skipped, the update pass runs without a spectrum (same bits as
            # the three passes on z - w = 0; the draw counter advances as usual, so the later minibatches are unchanged)
            self._grad(self.z, self.w, self.bits_mb, False, self.sptr, phases=4, spectrum=False, **gk)
            return self._prox(slot)
        # minibatch selection (device sampler, keyed by problem index) in parallel with the line pass
        self.ev_fork.record(self.stream)
        self.side.wait_event(self.ev_fork)
        self.check(self.lib.pnp_csmri_sel_sample(D.ptr(self.bits_mb), self.H, self.W, self.nb, D.ptr(self.support), D.ptr(self.m0),
                                                 self.sup_stride, self.B, self.seed & 0xffffffff, D.ptr(draws), None, 0,
                                                 self.side.cuda_stream))
        self.ev_join.record(self.side)
        self._grad(self.z, self.w, self.bits_mb, False, self.sptr, phases=1, **gk)
        self.stream.wait_event(self.ev_join)
        self._grad(self.z, self.w, self.bits_mb, False, self.sptr, phases=6, **gk)
        self._prox(slot)

    def _prox(self, slot):
        if self.fused_prox:
            rc = self.lib.pnp_prox_wavelet_fused(D.ptr(self.z), D.ptr(self.z), self.H, self.W, self.nb, D.ptr(self.sig_log),
                                                 self.sigma_modifier, 0.0, D.ptr(self.xrec), D.ptr(self.mse_log), D.ptr(slot), self.sptr)
            if rc == 0:
                self.check(self.lib.pnp_advance(D.ptr(self.counters), 3, self.sptr))
                return
            if rc != -4:
                self.check(rc)
            self.fused_prox = False
        self.check(self.lib.pnp_estimate_sigma(D.ptr(self.z), self.H, self.W, self.nb, D.ptr(self.sig_log), D.ptr(slot), self.sptr))
        self.check(self.lib.pnp_wavelet_denoise(D.ptr(self.z), D.ptr(self.z), self.H, self.W, self.nb, D.ptr(self.sig_log), 0.0,
                                                self.sigma_modifier, 0.0, D.ptr(self.xrec), D.ptr(self.mse_log), D.ptr(slot),
                                                self.sptr))
        self.check(self.lib.pnp_advance(D.ptr(self.counters), 3, self.sptr))

    def _capture(self):
        exec_ = C.c_void_p()
        self.stream.synchronize()
        self.check(self.lib.pnp_graph_begin(self.sptr))
        try:
            with torch.cuda.stream(self.stream):
                self._inner()
        finally:
            rc = self.lib.pnp_graph_end(self.sptr, C.byref(exec_))
        self.check(rc)
        self.graph = exec_

    def _capture_run(self, n_inner):
        """The WHOLE run (every snapshot and inner iteration) as one executable graph: a reused engine then costs one
        launch per batch instead of ~1.1 per iteration, which is what the host cores of an 8-rank sweep run out of."""
        exec_ = C.c_void_p()
        self.stream.synchronize()
        self.check(self.lib.pnp_graph_begin(self.sptr))
        try:
            with torch.cuda.stream(self.stream):
                done = 0
                while done < n_inner:
                    self._snapshot()
                    k = min(self.T2, n_inner - done)
                    for i in range(k):
                        self._inner(first_of_epoch=(i == 0))
                    done += k
        finally:
            rc = self.lib.pnp_graph_end(self.sptr, C.byref(exec_))
        self.check(rc)
        self._run_graph, self._run_graph_n = exec_, n_inner

    def run(self, n_inner):
        """n_inner inner iterations in total (snapshot every T2), nothing is read back."""
        if self.slots_used + n_inner > self.max_slots:
            raise ValueError('log capacity exceeded')
        if self.use_small:
            return self._run_small(n_inner)
        if self.lr_decay == 1.0 and self.whole_run_graph:
            if getattr(self, '_run_graph_n', None) != n_inner:
                if getattr(self, '_run_graph', None):
                    self.lib.pnp_graph_destroy(self._run_graph)
                self._capture_run(n_inner)
            with torch.cuda.stream(self.stream):
                self.check(self.lib.pnp_graph_launch(self._run_graph, self.sptr))
            self.outer += -(-n_inner // self.T2)
            self.slots_used += n_inner
            return
        with torch.cuda.stream(self.stream):
            if self.graph is None:
                self._capture()
            done = 0
            while done < n_inner:
                if self.lr_decay != 1.0:
                    self._decayed_step()
                self._snapshot()
                k = min(self.T2, n_inner - done)
                for _ in range(k):
                    self.check(self.lib.pnp_graph_launch(self.graph, self.sptr))
                done += k
                self.outer += 1
        self.slots_used += n_inner

    def _run_small(self, n_inner):
        """the whole run in one launch: every problem is a cluster of 8 CTAs that keeps its image, snapshot and snapshot
        gradient in shared memory (pnp_csmri_svrg_small)"""
        with torch.cuda.stream(self.stream):
            if self.lr_decay != 1.0:
                self._decayed_step()
            args = _lib.SvrgSmallArgs(
                H=self.H, W=self.W, batch=self.nb, z=D.ptr(self.z), xrec=D.ptr(self.xrec),
                Y1=D.ptr(self.Y1), Y2=D.ptr(self.Y2), Y1n=D.ptr(self.Y1n), Y2n=D.ptr(self.Y2n), bits_full=D.ptr(self.bits_full),
                support=D.ptr(self.support), m0=D.ptr(self.m0), support_img_stride=self.sup_stride,
                idx=None, idx_img_stride=0, idx_iter_stride=0, snap_scale_ptr=D.ptr(self.inv_m0), snap_scale=0.0,
                step=D.ptr(self.step), step_img_stride=1, sig_log=D.ptr(self.sig_log), mse_log=D.ptr(self.mse_log),
                slot=D.ptr(self.counters[0:1]), draw_counter=D.ptr(self.counters[2:3]),
                n_inner=int(n_inner), T2=self.T2, mini_batch_size=self.B, seed=self.seed & 0xffffffff,
                lr_decay=self.lr_decay, sigma_modifier=self.sigma_modifier, fallback_sigma=0.0, fallback_decay=1.0)
            self.check(self.lib.pnp_csmri_svrg_small(C.byref(args), self.sptr))
            self.check(self.lib.pnp_advance_by(D.ptr(self.counters), 3, int(n_inner), self.sptr))
        self.outer += -(-n_inner // self.T2)
        self.slots_used += n_inner

    def results(self, with_z=True):
        """-> dict(z [nb][N] float64 in the reference's raveled order, psnr [slots][nb], sigma_est)"""
        self.stream.synchronize()
        n = self.slots_used
        mse = self.mse_log[:n * self.nb].cpu().numpy().reshape(n, self.nb)
        sig = self.sig_log[:n * self.nb].cpu().numpy().reshape(n, self.nb) / self.W
        with np.errstate(divide='ignore'):
            psnr = np.around(10.0 * np.log10(self.data_range[None, :] ** 2 / (mse / self.N)), 2)
        z = (self.z.reshape(self.nb, self.W, self.H).transpose(1, 2).contiguous().cpu().numpy().astype(np.float64)
             if with_z else np.zeros((self.nb, 0)))
        with np.errstate(divide='ignore'):
            psnr0 = np.around(10.0 * np.log10(self.data_range ** 2 / (self.mse0.cpu().numpy() / self.N)), 2)
        return dict(z=z.reshape(self.nb, -1), psnr=psnr, psnr_init=psnr0, sigma_est=sig, m0=self.m0.cpu().numpy())

    def close(self):
        if self.graph:
            self.lib.pnp_graph_destroy(self.graph)
            self.graph = None
        if self._run_graph:
            self.lib.pnp_graph_destroy(self._run_graph)
            self._run_graph, self._run_graph_n = None, None
