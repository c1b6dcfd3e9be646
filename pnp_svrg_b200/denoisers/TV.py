"""TVDenoiser -- same signature as the reference's denoisers/TV.py:9-26.  As in the reference the
"TV" denoiser is skimage's wavelet BayesShrink (per-column multi-level Haar soft threshold);
the GPU kernel is csrc/prox.cuh::k_haar_bayes.

Additive mode (SURVEY section 8(a')): ``method='chambolle'`` is a true total-variation prox, Chambolle's
dual projection with skimage's ``denoise_tv_chambolle`` arithmetic and a fixed iteration count
(csrc/tv_chambolle.cuh).  ``weight`` is its regularisation weight; when ``weight`` is None it is
``sigma_est * sigma_modifier`` (the noise level the loops estimate), else ``denoise_strength * decay**t``."""
import numpy as np
import torch

from .. import _lib, device as D
from .denoiser import Denoise


class TVDenoiser(Denoise):
    def __init__(self, multi=True, rescale_sigma=True, decay=1, denoise_strength=0, sigma_modifier=1, *,
                 method='wavelet', weight=None, n_iter=20):
        super().__init__()
        if method not in ('wavelet', 'chambolle'):
            raise ValueError("method must be 'wavelet' (the reference's behaviour) or 'chambolle'")
        if method == 'chambolle' and int(n_iter) < 1:
            raise ValueError('n_iter must be >= 1')
        self.method, self.weight, self.n_iter = method, weight, int(n_iter)
        self._work = None
        if not multi:
            raise NotImplementedError('only multi=True (the reference default, per-column transform) is built')
        self.multi = multi
        self.rescale_sigma = rescale_sigma      # no effect on float input, as in skimage
        self.denoise_strength = denoise_strength
        self.sigma_modifier = sigma_modifier
        self.decay = decay

    def _dev_chambolle(self, ctx):
        self.t += 1
        n = ctx.H * ctx.W
        if self._work is None or self._work.numel() != 4 * n or self._work.device != ctx.z_in.device:
            self._work = torch.empty(4 * n, dtype=torch.float32, device=ctx.z_in.device)
            self._tmp = torch.empty(n, dtype=torch.float32, device=ctx.z_in.device)
        alias = ctx.z_out.data_ptr() == ctx.z_in.data_ptr()
        dst = self._tmp if alias else ctx.z_out
        fixed = float(self.weight) if self.weight is not None else 0.0
        from_est = self.weight is None
        _lib.check(_lib.load().pnp_tv_chambolle(
            D.ptr(ctx.z_in), D.ptr(dst), ctx.H, ctx.W, 1, fixed, D.ptr(ctx.sig_log) if from_est else None,
            float(self.sigma_modifier),
            float(ctx.sigma_est * self.sigma_modifier) if (from_est and ctx.sig_log is None and ctx.sigma_est > 0)
            else float(self.denoise_strength * self.decay ** self.t),
            self.n_iter, D.ptr(self._work), D.ptr(ctx.xrec), D.ptr(ctx.mse_log), D.ptr(ctx.slot), D.stream()))
        if alias:
            ctx.z_out.view(-1).copy_(dst)

    def _dev_denoise(self, ctx):
        if self.method == 'chambolle':
            return self._dev_chambolle(ctx)
        self.t += 1
        fallback = float(self.denoise_strength * self.decay ** self.t)
        _lib.check(_lib.load().pnp_wavelet_denoise(
            D.ptr(ctx.z_in), D.ptr(ctx.z_out), ctx.H, ctx.W, 1, D.ptr(ctx.sig_log), float(ctx.sigma_est),
            float(self.sigma_modifier), fallback, D.ptr(ctx.xrec), D.ptr(ctx.mse_log), D.ptr(ctx.slot), D.stream()))

    def _dev_prox_fused(self, ctx):
        """estimate_sigma + denoise + PSNR in one cooperative launch (ctx.sig_log must be the slot array the
        estimate is accumulated into).  Returns False when the image does not fit the SMs' shared memory."""
        if self.method != 'wavelet':
            return False
        rc = _lib.load().pnp_prox_wavelet_fused(
            D.ptr(ctx.z_in), D.ptr(ctx.z_out), ctx.H, ctx.W, 1, D.ptr(ctx.sig_log), float(self.sigma_modifier),
            float(self.denoise_strength * self.decay ** (self.t + 1)), D.ptr(ctx.xrec), D.ptr(ctx.mse_log), D.ptr(ctx.slot),
            D.stream())
        if rc == -4:            # PNP_ERR_UNSUPPORTED
            return False
        _lib.check(rc)
        self.t += 1
        return True

    def denoise(self, noisy, sigma_est=0):
        from ..engine import ProxCtx
        noisy = np.asarray(noisy)
        H, W = noisy.shape
        dev = D.require_cuda()
        z = D.to_lines(noisy, H, W, dev)
        out = torch.empty_like(z)
        s = float(sigma_est)
        self._dev_denoise(ProxCtx(z, out, H, W, sigma_est=s if s == s else 0.0))
        return D.from_lines(out, H, W).reshape(H, W)
