"""Phase retrieval (dense real Gaussian A, amplitude loss) -- same public surface as the
reference's problems/PR.py:12-87, gradients on the GPU through pnp_pr_grad (csrc/pr.cuh).

Additive mode (SURVEY section 8(a'), BASELINE config 3): ``model='cdp'`` is coded-diffraction-pattern phase
retrieval with the intensity loss, ``A_l x = fft2(d_l o x)/sqrt(N)``, ``d_l = 1j**codes[l]``, ``n_masks``
patterns, ``M = n_masks*N`` measurements ``y = |A x|^2``, gradient ``Re(A^H((|Az|^2 - y) o Az))`` through
pnp_cdp_grad (csrc/cdp.cuh).  The default stays the reference's dense model."""
import ctypes as C

import os

import numpy as np
import torch

from .. import _lib, device as D
from .problem import Problem


def shard_block(M, rank, world):
    """rows [lo, hi) of A owned by `rank`: contiguous, disjoint, covering [0, M) (the last rank takes the remainder)"""
    per = M // world
    lo = rank * per
    return lo, (M if rank == world - 1 else lo + per)


class PhaseRetrieval(Problem):
    def __init__(self, img_path=None, H=256, W=256, num_meas=-1, snr=None, sigma=None, *, image=None,
                 model='dense', n_masks=4, loss=None, shard=None):
        super().__init__(img_path, H, W, image=image)
        self.pname = 'pr'
        if model not in ('dense', 'cdp'):
            raise ValueError("model must be 'dense' (the reference's) or 'cdp'")
        self.model = model
        # shard = (rank, world): the FULL gradient (problems/PR.py:75-79, the M*N*4-byte stream that dominates an SVRG
        # epoch) is cut into contiguous row blocks of A, one per rank; every rank computes the partial sum of its rows
        # (already divided by the global M) and the partial gradients are summed by one all-reduce of 4N bytes
        # (``_snapshot_allreduce``, used by pnp_svrg's snapshot and by grad_full).  Minibatch gradients are replicated.
        if shard is not None:
            if model != 'dense':
                raise NotImplementedError('shard= is built for the dense model')
            rank, world = int(shard[0]), int(shard[1])
            if not (0 <= rank < world):
                raise ValueError('shard = (rank, world) with 0 <= rank < world')
            shard = (rank, world)
        self.shard = shard
        if model == 'cdp':
            if loss not in (None, 'intensity'):
                raise NotImplementedError("the coded-diffraction model is built with loss='intensity'")
            self._init_cdp(int(n_masks), snr, sigma)
            return
        if loss not in (None, 'amplitude'):
            raise NotImplementedError("the dense model is the reference's amplitude loss")
        self.M = num_meas
        self.snr = snr
        self.sigma = sigma
        self.A = np.random.randn(self.M, self.N)                # problems/PR.py:26
        self.Y0 = self.forward_model(self.X).ravel()
        self.set_snr_sigma()
        noises = np.random.normal(0, self.sigma, self.Y0.shape)
        self.Y = self.Y0 + noises
        self.SNR = self.get_snr_from_sigma
        self.spec_init()
        self.Xinit = (self.Xinit - self.Xinit.min()) / (self.Xinit.max() - self.Xinit.min())
        self._upload()

    # ---- construction (host float64, one-off) -----------------------------------------------
    def spec_init(self):
        """problems/PR.py:50-63 power iteration on D = A^T diag(Y) A / M, applied matrix-free
        (the reference materialises the N x N matrix)."""
        nrm = np.linalg.norm(self.X)
        A, Y = self.A, self.Y

        def apply(v):
            return A.T.dot(Y * A.dot(v)) / self.M
        m, mold = 1, 2
        cur, old = 2 * np.ones(self.N), np.ones(self.N)
        tol = 1e-5
        while abs(m - mold) > tol and np.linalg.norm(cur - old) > tol:
            mold, old = m, cur
            cur = apply(cur)
            m = np.max(cur)
            cur = cur / m
        self.Xinit = np.sqrt(m) * cur / np.linalg.norm(cur) * nrm

    def forward_model(self, w):
        if self.model == 'cdp':
            return np.abs(self._A_cdp(w)) ** 2
        return np.absolute(self.A.dot(np.asarray(w, dtype=np.float64).ravel()))

    def f(self, w):
        if self.model == 'cdp':
            return np.linalg.norm(self.Y - self.forward_model(w).ravel()) ** 2 / 4 / self.M
        return np.linalg.norm(self.Y - self.forward_model(w)) ** 2 / 2 / self.M

    def _upload(self):
        dev = self._device
        # columns of A permuted to the line layout: A_dev[m][c*H + r] = A[m][r*W + c]
        At = self.A.reshape(self.M, self.H, self.W).transpose(0, 2, 1).reshape(self.M, self.N)
        self._A = torch.from_numpy(np.ascontiguousarray(At, dtype=np.float32)).to(dev)
        self._y = torch.from_numpy(self.Y.astype(np.float32)).to(dev)
        self._r = torch.empty(self.M, dtype=torch.float32, device=dev)
        # row-chunk partial sums of the transposed product (pnp_pr_grad_args.partial): without them the column pass of a
        # 64 x 64 image is 4 CTAs (measured 492 us for a 33.6 MB matrix; 13.8 us for the row pass over the same bytes)
        self._chunks = 64
        self._partial = torch.empty(self._chunks * self.N, dtype=torch.float32, device=dev)
        self._shard_sel = None
        if self.shard is not None:
            lo, hi = shard_block(int(self.M), *self.shard)
            self._shard_block = (lo, hi)
            self._shard_sel = torch.arange(lo, hi, dtype=torch.int32, device=dev)

    def _snapshot_allreduce(self, mu):
        """sum the per-rank partial full gradients (NCCL all-reduce of 4N bytes)"""
        if self.shard is not None and self.shard[1] > 1:
            import torch.distributed as dist
            dist.all_reduce(mu, op=dist.ReduceOp.SUM)

    # ---- device protocol -----------------------------------------------------------------------
    def _dev_new_sel(self, count=0):
        return torch.zeros(max(int(count), 1), dtype=torch.int32, device=self._device)

    def _dev_set_sel(self, sel, idx_dev, count, cursor=None, stride=0, clear=True):
        _lib.check(_lib.load().pnp_copy_f32(D.ptr(sel), D.ptr(idx_dev), int(count), D.stream()))

    def _dev_sample_sel(self, sel, count, seed, counter=None, idx_out=None, clear=True):
        _lib.check(_lib.load().pnp_sample_indices(D.ptr(sel), int(self.M), int(count), int(seed) & 0xffffffff,
                                                  D.ptr(counter), D.stream()))

    def _dev_grad(self, a, b=None, sel=None, with_y=True, gscale=1.0, gscale_ptr=None, step=0.0, step_ptr=None,
                  g_out=None, vadd=None, v_out=None, z_in=None, z_out=None, phases=0, clear_sel=False, partial_ok=False):
        """g = [A_sel^T r(a)] - [A_sel^T r(b)] (two-point form when b is given), scaled by gscale."""
        if self.model == 'cdp':
            return self._dev_grad_cdp(a, b, sel, gscale, step, step_ptr, g_out, vadd, v_out, z_in, z_out)
        if sel is None and self._shard_sel is not None:
            if not partial_ok:
                raise NotImplementedError('PhaseRetrieval(shard=...): the full gradient of a sharded problem is a partial sum; '
                                          'only pnp_svrg (which all-reduces its snapshot) and grad_full() handle it')
            if self._shard_sel.numel() == 0:
                if g_out is not None:
                    g_out.zero_()
                return
            sel = self._shard_sel
        args = _lib.PrGradArgs(
            A=D.ptr(self._A), n=self.N, M=int(self.M), z=D.ptr(a), w=D.ptr(b), y=D.ptr(self._y), rows=D.ptr(sel),
            count=0 if sel is None else int(sel.numel()), cursor=None, r=D.ptr(self._r), gscale=float(gscale),
            step=float(step), step_ptr=D.ptr(step_ptr), g_out=D.ptr(g_out), vadd=D.ptr(vadd), v_out=D.ptr(v_out),
            z_in=D.ptr(z_in), z_out=D.ptr(z_out), partial=D.ptr(self._partial), partial_chunks=self._chunks)
        _lib.check(_lib.load().pnp_pr_grad(C.byref(args), D.stream()))

    # ---- coded diffraction patterns (additive) ------------------------------------------------------
    def _init_cdp(self, n_masks, snr, sigma):
        if n_masks < 1:
            raise ValueError('n_masks must be >= 1')
        self.L = n_masks
        self.M = self.L * self.N
        self.snr, self.sigma = snr, sigma
        self.codes = np.random.randint(0, 4, size=(self.L, self.H, self.W))           # d = 1j ** codes
        self._d = 1j ** self.codes
        self.Y0 = self.forward_model(self.X).ravel()
        self.set_snr_sigma()
        self.Y = self.Y0 + np.random.normal(0, self.sigma, self.Y0.shape)
        self.SNR = self.get_snr_from_sigma
        self._spec_init_cdp()
        self.Xinit = (self.Xinit - self.Xinit.min()) / (self.Xinit.max() - self.Xinit.min())
        dev = self._device
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.reshape(self.L, self.H, self.W).transpose(0, 2, 1), dtype=dt)).to(dev)
        self._codes = t(self.codes, np.int8)
        self._y = t(self.Y, np.float32)
        self._mask = torch.zeros(self.M, dtype=torch.uint8, device=dev)
        self._S = torch.empty(2 * self.M, dtype=torch.float32, device=dev)
        self._S2 = torch.empty(2 * self.M, dtype=torch.float32, device=dev)      # second point (SVRG / SARAH differences)
        self._acc = torch.empty(self.N, dtype=torch.float32, device=dev)

    def _A_cdp(self, w):
        return np.fft.fft2(self._d * np.asarray(w, dtype=np.float64).reshape(1, self.H, self.W)) / np.sqrt(self.N)

    def _AH_cdp(self, r):
        return (np.conj(self._d) * np.fft.ifft2(r) * np.sqrt(self.N)).sum(0)

    def _spec_init_cdp(self):
        """Spectral initialisation, the reference's power iteration (problems/PR.py:50-63) on
        D = A^H diag(Y) A / M applied matrix-free with FFTs."""
        nrm = np.linalg.norm(self.X)
        Y = self.Y.reshape(self.L, self.H, self.W)
        m, mold = 1, 2
        cur, old = 2 * np.ones(self.N), np.ones(self.N)
        tol = 1e-5
        it = 0
        while abs(m - mold) > tol and np.linalg.norm(cur - old) > tol and it < 500:
            mold, old = m, cur
            cur = np.real(self._AH_cdp(Y * self._A_cdp(cur))).ravel() / self.M
            m = np.max(cur)
            cur = cur / m
            it += 1
        self.Xinit = np.sqrt(abs(m)) * cur / np.linalg.norm(cur) * nrm

    def _dev_grad_cdp(self, a, b, sel, gscale, step, step_ptr, g_out, vadd, v_out, z_in, z_out):
        args = _lib.CdpGradArgs(
            H=self.H, W=self.W, L=self.L, codes=D.ptr(self._codes), y=D.ptr(self._y), z=D.ptr(a), w=D.ptr(b),
            sel_idx=D.ptr(sel), count=0 if sel is None else int(sel.numel()), cursor=None, mask=D.ptr(self._mask),
            S=D.ptr(self._S), acc=D.ptr(self._acc), gscale=float(gscale), step=float(step), step_ptr=D.ptr(step_ptr),
            g_out=D.ptr(g_out), vadd=D.ptr(vadd), v_out=D.ptr(v_out), z_in=D.ptr(z_in), z_out=D.ptr(z_out),
            S2=D.ptr(self._S2) if (b is not None and os.environ.get('PNP_CDP_TWO_PASS', '0') != '1') else None)
        _lib.check(_lib.load().pnp_cdp_grad(C.byref(args), D.stream()))

    # ---- reference API -----------------------------------------------------------------------
    def grad_full(self, z):
        """problems/PR.py:75-79."""
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        self._dev_grad(zl, gscale=1.0 / self.M, g_out=g, partial_ok=True)
        if getattr(self, 'shard', None) is not None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and self.shard[1] > 1:
                self._snapshot_allreduce(g)        # every rank returns the FULL gradient
            # (without a process group the caller gets this shard's partial gradient: the single-process shard tests)
        return D.from_lines(g, self.H, self.W)

    def grad_stoch(self, z, mb):
        """problems/PR.py:81-87."""
        idx = self._indices_of(mb)
        if idx.size == 0:
            return np.zeros(self.N)
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        self._dev_grad(zl, sel=torch.from_numpy(idx).to(self._device), g_out=g)
        return D.from_lines(g, self.H, self.W)
