"""Problem base class -- same public surface as the reference's problems/problem.py:8-175.

State that the PnP loop touches lives on the GPU in float32 in the transposed "line" layout
(see csrc/csmri.cuh); the NumPy attributes the reference exposes (``X``, ``Xrec``, ``Xinit``,
``Y`` ...) are kept on the host in float64 for scripts that read them.
"""
import numpy as np
import torch

from .. import _lib, device as D


class MiniBatch(np.ndarray):
    """The 0/1 integer array the reference's select_mb returns (problems/problem.py:110-117,
    CSMRI.py:66-74), carrying the drawn positions so the device path never has to scan it."""

    def __new__(cls, dense, indices):
        obj = np.asarray(dense).view(cls)
        obj.indices = np.asarray(indices, dtype=np.int32)
        return obj

    def __array_finalize__(self, obj):
        # survives reshape/ravel and element-wise products with the mask; dropped on slicing
        same = obj is not None and getattr(obj, 'size', None) == self.size
        self.indices = getattr(obj, 'indices', None) if same else None


def load_image(img_path, image, H, W):
    """problems/problem.py:16-25: PIL open -> resize((H, W)) -> min-max normalise to [0, 1].
    ``image`` (additive kwarg) supplies an already loaded 2-D array instead of a path."""
    from PIL import Image
    if img_path is not None:
        tmp = np.array(Image.open(img_path).resize((H, W)))
    elif image is not None:
        tmp = np.asarray(image)
        if tmp.shape != (H, W):
            tmp = np.array(Image.fromarray(tmp).resize((H, W)))
    else:
        raise Exception('Need to pass in image path or image')
    return (tmp - np.min(tmp)) / (np.max(tmp) - np.min(tmp))


class Problem():
    def __init__(self, img_path, H, W, image=None):
        self.H = H
        self.W = W
        self.N = H * W
        self.M = self.N
        self.Xrec = load_image(img_path, image, H, W)
        self.X = self.Xrec.ravel()
        self.Xinit = np.empty_like(self.X)
        if (H & (H - 1)) or (W & (W - 1)) or not (32 <= H <= 4096) or not (32 <= W <= 4096):
            raise Exception('pnp_svrg_b200 kernels need H and W to be powers of two in [32, 4096] '
                            '(got %dx%d)' % (H, W))
        self._device = D.require_cuda()
        self._xrec_dev = D.to_lines(self.Xrec, H, W, self._device)
        # skimage.metrics.peak_signal_noise_ratio: float images use data_range 1 when min >= 0
        self._data_range = 1.0 if self.Xrec.min() >= 0 else 2.0
        self._scratch_d = torch.zeros(4, dtype=torch.float64, device=self._device)

    # ---- reference API ---------------------------------------------------------------------
    def get_item(self, key):
        return self.__dict__[key]

    def PSNR(self, w):
        """problems/problem.py:33-35: round(10 log10(data_range^2 / MSE), 2) against Xrec."""
        z = D.to_lines(w, self.H, self.W, self._device)
        return self._psnr_from_sum(self._sq_err_dev(z))

    def _sq_err_dev(self, z_lines):
        self._scratch_d.zero_()
        _lib.check(_lib.load().pnp_sq_err(D.ptr(z_lines), D.ptr(self._xrec_dev), self.N, 1,
                                          D.ptr(self._scratch_d), None, D.stream()))
        return float(self._scratch_d[0].item())

    def _psnr_from_sum(self, sq_sum):
        with np.errstate(divide='ignore'):
            return np.around(10.0 * np.log10(self._data_range ** 2 / (np.asarray(sq_sum, dtype=np.float64) / self.N)), decimals=2)[()]

    def set_snr_sigma(self):
        if self.snr is not None and self.sigma is None:
            self.sigma = self.get_sigma_from_snr()
        elif self.sigma is not None and self.snr is None:
            self.snr = self.get_snr_from_sigma()
        elif self.snr is None and self.sigma is None:
            self.sigma = 0
            self.snr = 10e9
        else:
            raise Exception('Please specify either sigma (sigma) or signal-to-noise ratio (snr).')

    def get_snr_from_sigma(self):
        # norm (not norm**2), exactly as the reference computes it (problems/problem.py:48-56)
        if self.sigma > 0:
            return 10 * np.log10(np.linalg.norm(self.Y0.ravel()) / self.sigma ** 2 / self.H / self.W)
        elif self.sigma == 0:
            return 10e9
        raise Exception('Sigma cannot be negative.')

    def get_sigma_from_snr(self):
        return np.sqrt(np.linalg.norm(self.Y0.ravel()) / 10 ** (self.snr / 10) / self.H / self.W)

    def display(self, color_map='gray', show_measurements=False, save_results=False, save_dir='figures/',
                show_figs=False):
        """Plotting passthrough (out of the hot path): same figures as problems/problem.py:64-108."""
        self.color_map = color_map
        import matplotlib.pyplot as plt
        base = None
        if save_results:
            from datetime import datetime
            import os
            base = save_dir + self.pname + '/' + datetime.now().strftime('%y-%m-%d-%H-%M') + '/'
            self.prob_dir = base
            os.makedirs(base, exist_ok=True)
        panels = [('Original Image', self.Xrec, 'original.eps'),
                  ('Initialization', self.Xinit.reshape(self.H, self.W), 'initialization.eps')]
        if show_measurements:
            panels.append(('Measurements', np.real(self.Y).reshape(self.lrH, self.lrW), 'measurements.eps'))
        for title, img, fname in panels:
            fig = plt.figure(figsize=(6, 6))
            plt.imshow(img, cmap=color_map, vmin=0, vmax=1)
            plt.title(title)
            plt.xticks([])
            plt.yticks([])
            if base:
                fig.savefig(base + fname, transparent=True, bbox_inches='tight', pad_inches=0)
            if show_figs:
                plt.show()

    def _draw_indices(self, size):
        """Positions of a minibatch, consuming np.random exactly like problems/problem.py:115."""
        if size > self.M:
            print('MB size is too big: ', size, ' > ', self.M)
        return np.random.choice(self.M, size, replace=False)

    def select_mb(self, size):
        locs = self._draw_indices(size)
        batch = np.zeros(self.M)
        batch[locs] = 1
        return MiniBatch(batch.astype(int), locs)

    def _indices_of(self, mb):
        idx = getattr(mb, 'indices', None)
        if idx is None:
            idx = np.flatnonzero(np.asarray(mb).ravel())
        return np.ascontiguousarray(idx, dtype=np.int32)

    def f(self, z):
        raise NotImplementedError('Need to implement f() method')

    def grad_full(self, z):
        raise NotImplementedError('Need to implement full_grad() method')

    def grad_stoch(self, z, mb_indices):
        raise NotImplementedError('Need to implement stoch_grad() method')

    def grad_full_check(self):
        """problems/problem.py:131-155 finite-difference check (O(N) evaluations of f)."""
        w = np.random.uniform(0.0, 1.0, self.N)
        delta = np.zeros(self.N)
        grad = np.zeros(self.N)
        eps = 1e-6
        f0 = self.f(w)
        for i in range(self.N):
            delta[i] = eps
            grad[i] = (self.f(w + delta) - f0) / eps
            delta[i] = 0
        grad_comp = np.asarray(self.grad_full(w)).ravel()
        ok = np.linalg.norm(grad - grad_comp) <= 1e-4
        print('Full Grad check succeeded!' if ok else 'Full Grad check failed!')
        return bool(ok)

    def grad_stoch_check(self):
        """problems/problem.py:157-175: sum_i grad_stoch(w, e_i) / M == grad_full(w)."""
        w = np.random.uniform(0.0, 1.0, self.N)
        full = np.asarray(self.grad_full(w)).ravel()
        acc = np.zeros(self.N)
        for i in range(self.M):
            mb = np.zeros(self.M, dtype=int)
            mb[i] = 1
            acc += np.asarray(self.grad_stoch(w, mb)).ravel()
        ok = np.linalg.norm(full - acc / self.M) <= 1e-6
        print('Stoch Grad check succeeded!' if ok else 'Stoch Grad check failed!')
        return bool(ok)
