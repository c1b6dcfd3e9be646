"""Deblur + super-resolution -- same public surface as the reference's problems/DeblurSR.py:16-147,
gradients on the GPU through pnp_deblur_grad (csrc/deblur.cuh)."""
import ctypes as C

import numpy as np
import torch
from PIL import Image

from .. import _lib, device as D
from .problem import Problem

eps = 1e-10


class _HostAdjoint:
    def __init__(self, op):
        self._op = op

    def __mul__(self, y):
        return self._op._rmatvec(np.asarray(y, dtype=np.float64))


class _HostSampler:
    """Host float64 twin of the sampling operator (pylops Identity / Bilinear in the reference,
    problems/DeblurSR.py:95-108); only used at construction (Y0) and exposed as ``problem.Bop``."""

    def __init__(self, dims, rows=None, cols=None):
        self.dims = dims
        self.identity = rows is None
        if not self.identity:
            self.t = np.floor(rows).astype(np.int64)
            self.l = np.floor(cols).astype(np.int64)
            self.wr = rows - self.t
            self.wc = cols - self.l

    @property
    def H(self):
        return _HostAdjoint(self)

    def __mul__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if self.identity:
            return x.ravel()
        g = x.reshape(self.dims)
        t, l, wr, wc = self.t, self.l, self.wr, self.wc
        return (g[t, l] * (1 - wr) * (1 - wc) + g[t, l + 1] * (1 - wr) * wc
                + g[t + 1, l] * wr * (1 - wc) + g[t + 1, l + 1] * wr * wc)

    def _rmatvec(self, y):
        if self.identity:
            return y.ravel()
        out = np.zeros(self.dims)
        t, l, wr, wc = self.t, self.l, self.wr, self.wc
        for dt, dl, w in ((0, 0, (1 - wr) * (1 - wc)), (0, 1, (1 - wr) * wc), (1, 0, wr * (1 - wc)), (1, 1, wr * wc)):
            np.add.at(out, (t + dt, l + dl), y * w)
        return out.ravel()


class Deblur(Problem):
    def __init__(self, img_path=None, H=64, W=64, kernel_path=None, kernel=None, scale_percent=50,
                 snr=None, sigma=None, *, image=None, conv='auto'):
        super().__init__(img_path, H, W, image=image)
        if conv not in ('auto', 'fft', 'direct'):
            raise ValueError("conv must be 'auto', 'fft' or 'direct'")
        self.conv = conv          # additive: 'direct' = tap sums for kernels with <= 16 non-zero entries (SURVEY 8(a'))
        self.pname = 'deblur'
        self.scale_percent = scale_percent
        self.snr = snr
        self.sigma = sigma
        self.kernel_path = kernel_path
        self.kernel = kernel
        if kernel_path is None and kernel is None:
            raise Exception('Need to pass in kernel path or kernel as image')
        self._load_kernel()
        self.lrH = int(self.H * scale_percent / 100)
        self.lrW = int(self.W * scale_percent / 100)
        self.M = self.lrH * self.lrW
        self._generate_bop()
        self.Y0 = self.forward_model(self.X)
        self.set_snr_sigma()
        noises = np.random.normal(0, self.sigma, self.Y0.shape)
        self.Y = self.Y0 + noises
        self.Xinit = np.random.uniform(0.0, 1.0, self.N)
        self._upload()

    # ---- construction (host, one-off) --------------------------------------------------------
    def _load_kernel(self):
        # problems/DeblurSR.py:72-93
        if self.kernel_path is not None:
            self.B = np.array(Image.open(self.kernel_path).resize((self.H, self.W)))
        elif isinstance(self.kernel, str) and self.kernel == "Identity":
            self.B = np.zeros(self.N)
            self.B[0] = 1
        elif isinstance(self.kernel, str) and self.kernel == "Minimal":
            self.B = np.zeros((self.H, self.W))
            self.B[0, 0] = 1
            for div in (2, 3, 4):
                self.B[self.H // 2, self.H // div] = 1
            self.B /= 4
        elif self.kernel is not None:
            self.B = np.asarray(self.kernel)
        else:
            raise Exception('Need to pass in blur kernel path or kernel')
        self.B = self.B.ravel() / self.N
        if self.B.size != self.N:
            raise Exception('kernel must have H*W entries')

    def _generate_bop(self):
        if self.scale_percent == 100:
            self.Bop = _HostSampler((self.H, self.W))
        else:
            ptsH = np.linspace(eps, self.H - (1 + eps), self.lrH)
            ptsW = np.linspace(eps, self.W - (1 + eps), self.lrW)
            meshW, meshH = np.meshgrid(ptsH, ptsW)
            rows, cols = meshH.ravel(), meshW.ravel()
            if np.unique(np.vstack([rows, cols]), axis=1).shape[1] != rows.size:
                raise ValueError('repeated values in iava array')
            self.Bop = _HostSampler((self.H, self.W), rows, cols)

    def fft_blur(self, M1, M2):
        # problems/DeblurSR.py:119-120 (host float64; the GPU path is pnp_deblur_grad)
        return np.real(np.fft.ifft(np.fft.fft(np.ravel(M1)) * np.fft.fft(np.ravel(M2)))) * np.sqrt(self.N)

    def fft_deblur(self, M1, M2):
        return np.real(np.fft.ifft(np.fft.fft(np.ravel(M1)) / np.fft.fft(np.ravel(M2))))

    def forward_model(self, w):
        return self.Bop * self.fft_blur(w, self.B)

    def f(self, w):
        return np.linalg.norm(self.Y - self.forward_model(w)) ** 2 / 2 / self.M

    def _upload(self):
        H, W, N, dev = self.H, self.W, self.N, self._device
        hp = H // 2
        F = np.fft.fft(self.B) * np.sqrt(N)                     # F[k1 + H*k2]
        Bf = F.reshape(W, H).T[:hp + 1]                         # [k1][k2], row hp = Nyquist
        self._Bf = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(Bf).astype(np.complex64)).to(dev)).contiguous()
        twn = np.exp(-2j * np.pi * np.arange(W) / N)
        self._twn = torch.view_as_real(torch.from_numpy(twn.astype(np.complex64)).to(dev)).contiguous()
        self._y = torch.from_numpy(self.Y.astype(np.float32)).to(dev)
        self._identity = bool(self.Bop.identity)
        if self._identity:
            self._tl = self._wts = None
        else:
            tl = np.stack([self.Bop.t, self.Bop.l], axis=1).astype(np.int32)
            wts = np.stack([self.Bop.wr, self.Bop.wc], axis=1).astype(np.float32)
            self._tl = torch.from_numpy(np.ascontiguousarray(tl)).to(dev)
            self._wts = torch.from_numpy(np.ascontiguousarray(wts)).to(dev)
        taps = np.flatnonzero(self.B)
        if self.conv == 'direct' and taps.size > 16:
            raise ValueError("conv='direct' needs a blur kernel with at most 16 non-zero entries (this one has %d)" % taps.size)
        self._direct = self.conv != 'fft' and 0 < taps.size <= 16
        self._tap_pos = np.ascontiguousarray(taps, dtype=np.int32)
        self._tap_w = np.ascontiguousarray(self.B[taps] * np.sqrt(N), dtype=np.float32)
        self._S = torch.empty(N, dtype=torch.float32, device=dev)
        self._blurred = torch.empty(N, dtype=torch.float32, device=dev)
        self._up = torch.empty(N, dtype=torch.float32, device=dev)

    # ---- device protocol -----------------------------------------------------------------------
    def _dev_new_sel(self, count=0):
        return torch.zeros(max(int(count), 1), dtype=torch.int32, device=self._device)

    def _dev_set_sel(self, sel, idx_dev, count, cursor=None, stride=0, clear=True):
        _lib.check(_lib.load().pnp_copy_f32(D.ptr(sel), D.ptr(idx_dev), int(count), D.stream()))

    def _dev_sample_sel(self, sel, count, seed, counter=None, idx_out=None, clear=True):
        _lib.check(_lib.load().pnp_sample_indices(D.ptr(sel), int(self.M), int(count), int(seed) & 0xffffffff,
                                                  D.ptr(counter), D.stream()))

    def _dev_grad(self, a, b=None, sel=None, with_y=True, gscale=1.0, gscale_ptr=None, step=0.0, step_ptr=None,
                  g_out=None, vadd=None, v_out=None, z_in=None, z_out=None, phases=0, clear_sel=False):
        args = _lib.DeblurGradArgs(
            H=self.H, W=self.W, batch=1, a=D.ptr(a), b=D.ptr(b), S=D.ptr(self._S), blurred=D.ptr(self._blurred),
            up=D.ptr(self._up), Bf=D.ptr(self._Bf), twn=D.ptr(self._twn), y=D.ptr(self._y), tl=D.ptr(self._tl),
            wts=D.ptr(self._wts), identity=int(self._identity), M=int(self.M), sel=D.ptr(sel),
            count=0 if sel is None else int(sel.numel()), cursor=None, use_y=int(bool(with_y) and b is None),
            gscale=float(gscale), step=float(step), step_ptr=D.ptr(step_ptr), g_out=D.ptr(g_out), vadd=D.ptr(vadd),
            v_out=D.ptr(v_out), z_in=D.ptr(z_in), z_out=D.ptr(z_out),
            ntaps=int(self._tap_pos.size) if self._direct else 0,
            tap_pos=self._tap_pos.ctypes.data if self._direct else None,
            tap_w=self._tap_w.ctypes.data if self._direct else None)
        _lib.check(_lib.load().pnp_deblur_grad(C.byref(args), D.stream()))

    # ---- reference API -----------------------------------------------------------------------
    def grad_full(self, z):
        """problems/DeblurSR.py:126-132."""
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        self._dev_grad(zl, gscale=1.0 / self.M, g_out=g)
        return D.from_lines(g, self.H, self.W)

    def grad_stoch(self, z, mb):
        """problems/DeblurSR.py:135-147."""
        idx = self._indices_of(mb)
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        sel = torch.from_numpy(idx).to(self._device)
        if idx.size == 0:
            return np.zeros(self.N)
        self._dev_grad(zl, sel=sel, g_out=g)
        return D.from_lines(g, self.H, self.W)
