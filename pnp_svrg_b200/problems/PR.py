"""Phase retrieval (dense real Gaussian A, amplitude loss) -- same public surface as the
reference's problems/PR.py:12-87, gradients on the GPU through pnp_pr_grad (csrc/pr.cuh)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib, device as D
from .problem import Problem


class PhaseRetrieval(Problem):
    def __init__(self, img_path=None, H=256, W=256, num_meas=-1, snr=None, sigma=None, *, image=None):
        super().__init__(img_path, H, W, image=image)
        self.pname = 'pr'
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
        return np.absolute(self.A.dot(np.asarray(w, dtype=np.float64).ravel()))

    def f(self, w):
        return np.linalg.norm(self.Y - self.forward_model(w)) ** 2 / 2 / self.M

    def _upload(self):
        dev = self._device
        # columns of A permuted to the line layout: A_dev[m][c*H + r] = A[m][r*W + c]
        At = self.A.reshape(self.M, self.H, self.W).transpose(0, 2, 1).reshape(self.M, self.N)
        self._A = torch.from_numpy(np.ascontiguousarray(At, dtype=np.float32)).to(dev)
        self._y = torch.from_numpy(self.Y.astype(np.float32)).to(dev)
        self._r = torch.empty(self.M, dtype=torch.float32, device=dev)

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
        """g = [A_sel^T r(a)] - [A_sel^T r(b)] (two-point form when b is given), scaled by gscale."""
        args = _lib.PrGradArgs(
            A=D.ptr(self._A), n=self.N, M=int(self.M), z=D.ptr(a), w=D.ptr(b), y=D.ptr(self._y), rows=D.ptr(sel),
            count=0 if sel is None else int(sel.numel()), cursor=None, r=D.ptr(self._r), gscale=float(gscale),
            step=float(step), step_ptr=D.ptr(step_ptr), g_out=D.ptr(g_out), vadd=D.ptr(vadd), v_out=D.ptr(v_out),
            z_in=D.ptr(z_in), z_out=D.ptr(z_out))
        _lib.check(_lib.load().pnp_pr_grad(C.byref(args), D.stream()))

    # ---- reference API -----------------------------------------------------------------------
    def grad_full(self, z):
        """problems/PR.py:75-79."""
        zl = D.to_lines(z, self.H, self.W, self._device)
        g = torch.empty_like(zl)
        self._dev_grad(zl, gscale=1.0 / self.M, g_out=g)
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
