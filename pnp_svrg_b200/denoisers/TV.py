"""TVDenoiser -- same signature as the reference's denoisers/TV.py:9-26.  As in the reference the
"TV" denoiser is skimage's wavelet BayesShrink (per-column multi-level Haar soft threshold);
the GPU kernel is csrc/prox.cuh::k_haar_bayes."""
import numpy as np
import torch

from .. import _lib, device as D
from .denoiser import Denoise


class TVDenoiser(Denoise):
    def __init__(self, multi=True, rescale_sigma=True, decay=1, denoise_strength=0, sigma_modifier=1):
        super().__init__()
        if not multi:
            raise NotImplementedError('only multi=True (the reference default, per-column transform) is built')
        self.multi = multi
        self.rescale_sigma = rescale_sigma      # no effect on float input, as in skimage
        self.denoise_strength = denoise_strength
        self.sigma_modifier = sigma_modifier
        self.decay = decay

    def _dev_denoise(self, ctx):
        self.t += 1
        fallback = float(self.denoise_strength * self.decay ** self.t)
        _lib.check(_lib.load().pnp_wavelet_denoise(
            D.ptr(ctx.z_in), D.ptr(ctx.z_out), ctx.H, ctx.W, 1, D.ptr(ctx.sig_log), float(ctx.sigma_est),
            float(self.sigma_modifier), fallback, D.ptr(ctx.xrec), D.ptr(ctx.mse_log), D.ptr(ctx.slot), D.stream()))

    def _dev_prox_fused(self, ctx):
        """estimate_sigma + denoise + PSNR in one cooperative launch (ctx.sig_log must be the slot array the
        estimate is accumulated into).  Returns False when the image does not fit the SMs' shared memory."""
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
