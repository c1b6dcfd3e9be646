"""Compressive-sensing MRI -- same public surface as the reference's problems/CSMRI.py:11-89,
gradients on the GPU through pnp_csmri_grad (csrc/csmri.cuh)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib, device as D
from .problem import MiniBatch, Problem


def shard_rows(H, rank, world):
    """Band [lo, hi) of PACKED ky rows owned by `rank`: kyp = min(ky, H - ky) in [0, H/2); the Nyquist row ky = H/2
    rides in packed row 0 and belongs to rank 0.  The bands are contiguous, disjoint and cover [0, H/2) (the last rank
    takes the remainder).  A measurement at (ky, kx) and its Hermitian mirror land on the same rank, so the column
    pass of a rank only transforms its own packed rows (pnp_csmri_grad_args.row_lo / row_hi)."""
    hp = H // 2
    per = hp // world
    lo = rank * per
    return lo, (hp if rank == world - 1 else lo + per)


def packed_row_of(ky, H):
    """packed half-spectrum row of k-space row ky (see shard_rows)"""
    ky = np.asarray(ky)
    hp = H // 2
    return np.where(ky < hp, ky, np.where(ky == hp, 0, H - ky))


class CSMRI(Problem):
    def __init__(self, img_path=None, H=256, W=256, sample_prob=0.5, snr=None, sigma=None, *,
                 image=None, mask_type='bernoulli', shard=None):
        super().__init__(img_path, H, W, image=image)
        if H != W:
            raise Exception('CSMRI needs a square image (the reference applies an HxH DFT matrix on both sides)')
        self.pname = 'csmri'
        self.sample_prob = sample_prob
        self.snr = snr
        self.sigma = sigma
        self.mask_type = mask_type
        # shard = (rank, world): this process holds only the measurements whose ky row falls in its
        # contiguous row block; grad_full then returns a PARTIAL sum that the caller all-reduces
        # (the one real exchange step of the path).  Minibatch gradients and
        # everything else stay replicated.
        self.shard = shard

        self._generate_mask()
        self.Y0 = self.forward_model(self.X)
        self.set_snr_sigma()
        noises = np.random.normal(0, self.sigma, self.Y0.shape)
        self.Y = self.Y0 + np.multiply(self.mask, noises)
        self.SNR = self.get_snr_from_sigma
        x0 = np.absolute(np.fft.ifft2(self.Y)).ravel()
        self.Xinit = (x0 - np.min(x0)) / (np.max(x0) - np.min(x0))

        self.lrH, self.lrW = self.H, self.W
        self.M = self.N
        self.M0 = np.count_nonzero(self.mask)
        self._upload_measurements()

    # ---- construction (host, one-off; SURVEY section 8(f) rank 2 moves it to the device) ------
    def _generate_mask(self):
        if self.mask_type == 'bernoulli':      # problems/CSMRI.py:43-45
            self.mask = np.random.choice([0, 1], size=(self.H, self.W),
                                         p=[1 - self.sample_prob, self.sample_prob])
        elif self.mask_type == 'rows':         # additive mode named by the north star (not in the reference)
            rows = np.random.choice([0, 1], size=(self.H, 1), p=[1 - self.sample_prob, self.sample_prob])
            self.mask = np.repeat(rows, self.W, axis=1)
        else:
            raise Exception('unknown mask_type %r' % (self.mask_type,))

    def forward_model(self, w):
        """mask o fft2(w): the reference multiplies by a dense DFT matrix on both sides
        (problems/CSMRI.py:47-59), which equals fft2 to ~6e-9; the dense matrix is not built."""
        return np.multiply(self.mask, np.fft.fft2(np.asarray(w, dtype=np.float64).reshape(self.H, self.W)))

    def f(self, w):
        return np.linalg.norm(self.Y - self.forward_model(w)) ** 2 / 2 / self.M

    def _upload_measurements(self):
        H, W, hp = self.H, self.W, self.H // 2
        dev = self._device
        Ym = np.multiply(self.mask, self.Y)
        Ymir = np.conj(Ym[(-np.arange(H)) % H][:, (-np.arange(W)) % W])

        def up(a):
            t = torch.from_numpy(np.ascontiguousarray(a).astype(np.complex64)).to(dev)
            return torch.view_as_real(t).contiguous()
        self._Y1 = up(Ym[:hp, :])               # [kyp][kx]
        self._Y2 = up(Ymir[:hp, :])
        self._Y1n = up(Ym[hp, :])
        self._Y2n = up(Ymir[hp, :])
        self._support_host = np.flatnonzero(self.mask).astype(np.int32)
        self._support = torch.from_numpy(self._support_host).to(dev)
        self._m0_dev = torch.tensor([self.M0], dtype=torch.int32, device=dev)
        self._bits_full = torch.zeros(W * hp, dtype=torch.uint8, device=dev)
        if self.shard is None:
            self._dev_set_sel(self._bits_full, self._support, self.M0)
        else:
            lo, hi = shard_rows(self.H, *self.shard)
            rows = packed_row_of(self._support_host // W, self.H)
            own = np.ascontiguousarray(self._support_host[(rows >= lo) & (rows < hi)])
            self._shard_rows = (int(lo), int(hi))
            self._shard_count = int(own.size)
            own_dev = torch.from_numpy(own).to(dev)
            if own.size:
                self._dev_set_sel(self._bits_full, own_dev, own.size)
        self._S = torch.empty(self.N, dtype=torch.float32, device=dev)
        self._bits_tmp = torch.zeros(W * hp, dtype=torch.uint8, device=dev)

    # ---- device protocol used by pnp_svrg_b200.engine ------------------------------------------
    def _dev_new_sel(self, count=0):
        return torch.zeros(self.W * (self.H // 2), dtype=torch.uint8, device=self._device)

    def _dev_full_sel(self):
        return self._bits_full

    def _snapshot_allreduce(self, mu):
        """sum the per-rank partial snapshot gradients (NCCL all-reduce of 4N bytes over NVLink)"""
        if self.shard is not None and self.shard[1] > 1:
            import torch.distributed as dist
            dist.all_reduce(mu, op=dist.ReduceOp.SUM)

    def _dev_set_sel(self, sel, idx_dev, count, cursor=None, stride=0, clear=True):
        _lib.check(_lib.load().pnp_csmri_sel_from_indices(D.ptr(sel), self.H, self.W, 1, D.ptr(idx_dev), int(count),
                                                          int(stride), D.ptr(cursor), int(bool(clear)), D.stream()))

    def _dev_sample_sel(self, sel, count, seed, counter=None, idx_out=None, clear=True):
        if int(count) > self.M0:
            # the reference prints a warning and np.random.choice raises (problems/CSMRI.py:68-72); the device
            # sampler's cycle walk is only defined inside [0, M0)
            raise ValueError('Cannot take a larger sample (%d) than the %d sampled k-space positions' % (count, self.M0))
        _lib.check(_lib.load().pnp_csmri_sel_sample(D.ptr(sel), self.H, self.W, 1, D.ptr(self._support),
                                                    D.ptr(self._m0_dev), 0, int(count), int(seed) & 0xffffffff,
                                                    D.ptr(counter), D.ptr(idx_out), int(bool(clear)), D.stream()))

    _inpass_sel = True      # the forward line pass can build the minibatch selection itself (sel_job)
    _chain_ok = True        # the passes take chain=True (programmatic dependent launch)

    def _dev_grad(self, a, b=None, sel=None, with_y=True, gscale=1.0, gscale_ptr=None, step=0.0, step_ptr=None,
                  g_out=None, vadd=None, v_out=None, z_in=None, z_out=None, phases=0, clear_sel=False, sel_job=None,
                  chain=False, partial_ok=False):
        """g = Re(ifft2(sel o fft2(a - b) - Ysel)) * gscale ; v = g + vadd ; z_out = z_in - step*v.
        ``sel_job`` = dict(count, idx | None, cursor, seed, counter): pass 1 also builds the minibatch selection in
        ``sel`` (all zero on entry) -- from explicit positions ``idx`` or by the device sampler."""
        sj = {}
        if sel is None and self.shard is not None:
            # measurement shard: the full-mask gradient of this rank is a PARTIAL sum over its band of packed rows
            if not partial_ok:
                raise NotImplementedError('CSMRI(shard=...): the full gradient of a sharded problem is a partial sum; only '
                                          'pnp_svrg (which all-reduces its snapshot) and grad_full() handle it')
            sj = dict(row_lo=self._shard_rows[0], row_hi=self._shard_rows[1])
        if sel_job is not None:
            if sel is None:
                raise ValueError('sel_job needs the selection buffer it fills')
            if sel_job.get('idx') is None and int(sel_job['count']) > self.M0:
                raise ValueError('Cannot take a larger sample (%d) than the %d sampled k-space positions'
                                 % (sel_job['count'], self.M0))
            sj.update(sel_count=int(sel_job['count']), sel_idx=D.ptr(sel_job.get('idx')), sel_idx_img_stride=0,
                      sel_cursor=D.ptr(sel_job.get('cursor')), sel_support=D.ptr(self._support), sel_m0=D.ptr(self._m0_dev),
                      sel_support_img_stride=0, sel_seed=int(sel_job.get('seed', 0)) & 0xffffffff,
                      sel_counter=D.ptr(sel_job.get('counter')), sel_min_m0=int(self.M0))
        args = _lib.CsmriGradArgs(
            H=self.H, W=self.W, batch=1, a=D.ptr(a), b=D.ptr(b), S=D.ptr(self._S),
            bits=D.ptr(self._bits_full if sel is None else sel),
            Y1=D.ptr(self._Y1) if with_y else None, Y2=D.ptr(self._Y2) if with_y else None,
            Y1n=D.ptr(self._Y1n) if with_y else None, Y2n=D.ptr(self._Y2n) if with_y else None,
            gscale=float(gscale), gscale_ptr=D.ptr(gscale_ptr), step=float(step), step_ptr=D.ptr(step_ptr),
            g_out=D.ptr(g_out), vadd=D.ptr(vadd), v_out=D.ptr(v_out), z_in=D.ptr(z_in), z_out=D.ptr(z_out), phases=int(phases),
            clear_bits=int(bool(clear_sel) and sel is not None), flags=1 if chain else 0, **sj)
        _lib.check(_lib.load().pnp_csmri_grad(C.byref(args), D.stream()))

    def _dev_update_prox(self, gscale, step_ptr, vadd, z_in, z_out, sig_log, sigma_modifier, fallback_sigma, xrec, mse_log,
                         slot, advance=None, n_advance=0, barrier_ws=None, chain=False, zero_spectrum=False, next_pass=None):
        """Tail of an inner iteration in one cooperative launch, after ``_dev_grad(..., phases=3)`` left the masked
        spectrum in the scratch: inverse line pass + update + sigma estimate + wavelet prox + PSNR
        (pnp_csmri_update_prox).  Returns False when the image does not suit the resident-line kernel."""
        nxt = None
        if next_pass is not None:
            # the forward line pass (and the minibatch selection) of the NEXT inner iteration, fused into this launch:
            # next_pass = dict(w=snapshot, sel=selection bytes, job=Engine.sel_job()-style dict or None)
            job = next_pass.get('job')
            if job is not None and job.get('idx') is None and int(job['count']) > self.M0:
                raise ValueError('Cannot take a larger sample (%d) than the %d sampled k-space positions' % (job['count'], self.M0))
            nxt = _lib.CsmriNextPass(
                w=D.ptr(next_pass['w']), S_out=D.ptr(self._S), bits=D.ptr(next_pass.get('sel')),
                sel_count=0 if job is None else int(job['count']), sel_idx=None if job is None else D.ptr(job.get('idx')),
                sel_support=D.ptr(self._support), sel_m0=D.ptr(self._m0_dev),
                sel_seed=0 if job is None else int(job.get('seed', 0)) & 0xffffffff,
                sel_counter=None if job is None else D.ptr(job.get('counter')),
                sel_counter_add=0 if job is None else int(job.get('counter_add', 0)), sel_min_m0=int(self.M0))
        rc = _lib.load().pnp_csmri_update_prox_next(None if zero_spectrum else D.ptr(self._S), self.H, self.W, float(gscale), 0.0,
                                                    D.ptr(step_ptr), D.ptr(vadd), D.ptr(z_in), D.ptr(z_out), D.ptr(sig_log),
                                                    float(sigma_modifier), float(fallback_sigma), D.ptr(xrec), D.ptr(mse_log),
                                                    D.ptr(slot), D.ptr(advance), int(n_advance), D.ptr(barrier_ws), int(bool(chain)),
                                                    None if nxt is None else C.byref(nxt), D.stream())
        if rc == -4:            # PNP_ERR_UNSUPPORTED
            return False
        _lib.check(rc)
        return True

    # ---- reference API ---------------------------------------------------------------------
    def _draw_indices(self, size):
        # problems/CSMRI.py:66-74: uniform draw without replacement from the sampled support
        if size > self.M:
            print('MB size is too big: ', size, ' > ', self.M)
        return np.random.choice(self._support_host, size, replace=False)

    def select_mb(self, size):
        locs = self._draw_indices(size)
        batch = np.zeros(self.M)
        batch[locs] = 1
        return MiniBatch(batch.reshape(self.H, self.W).astype(int), locs)

    def grad_full(self, z):
        """problems/CSMRI.py:76-81."""
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        self._dev_grad(zl, gscale=1.0 / self.M0, g_out=g, partial_ok=True)
        if self.shard is not None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and self.shard[1] > 1:
                self._snapshot_allreduce(g)        # every rank returns the FULL gradient
            # (without a process group the caller gets this shard's partial gradient: the single-process shard tests)
        return D.from_lines(g, self.H, self.W)

    def grad_stoch(self, z, mb):
        """problems/CSMRI.py:83-89 (not divided by the batch size; positions outside the mask drop out)."""
        idx = self._indices_of(mb)
        idx = idx[self.mask.ravel()[idx] != 0]
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        idx_dev = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(self._device)
        self._dev_set_sel(self._bits_tmp, idx_dev, idx.size)
        self._dev_grad(zl, sel=self._bits_tmp, g_out=g)
        return D.from_lines(g, self.H, self.W)
