"""NLMDenoiser -- same signature as the reference's denoisers/NLM.py:9-27; classic (fast_mode=False)
non-local means on the GPU (csrc/nlm.cuh).

Documented deviation: the reference's denoise() reads ``self.sigma``, an attribute its __init__
never sets (AttributeError on first use).  The test used here is ``sigma_est > 0``, the convention of
the sibling TV / BM3D denoisers (denoisers/TV.py:23, denoisers/BM3D.py:22)."""
import numpy as np
import torch

from .. import _lib, device as D
from .denoiser import Denoise


class NLMDenoiser(Denoise):
    def __init__(self, decay=1, denoise_strength=0, patch_size=4, patch_distance=5, sigma_modifier=1,
                 fast_mode=False, multichannel=True):
        super().__init__()
        if fast_mode:
            raise NotImplementedError('only fast_mode=False (the reference default) is built')
        self.decay = decay
        self.denoise_strength = denoise_strength
        self.fast_mode = fast_mode
        self.sigma_modifier = sigma_modifier
        self.patch = dict(patch_size=patch_size, patch_distance=patch_distance, multichannel=multichannel)
        self._tmp = None

    def _dev_denoise(self, ctx):
        self.t += 1
        out = ctx.z_out
        alias = ctx.z_out.data_ptr() == ctx.z_in.data_ptr()
        if alias:                      # the kernel reads a halo: it cannot run in place
            if self._tmp is None or self._tmp.numel() != ctx.z_in.numel() or self._tmp.device != ctx.z_in.device:
                self._tmp = torch.empty_like(ctx.z_in)
            out = self._tmp
        _lib.check(_lib.load().pnp_nlm_denoise(
            D.ptr(ctx.z_in), D.ptr(out), ctx.H, ctx.W, 1, int(self.patch['patch_size']),
            int(self.patch['patch_distance']), D.ptr(ctx.sig_log), float(ctx.sigma_est), float(self.sigma_modifier),
            float(self.denoise_strength * self.decay ** self.t), D.ptr(ctx.xrec), D.ptr(ctx.mse_log), D.ptr(ctx.slot),
            D.stream()))
        if alias:
            _lib.check(_lib.load().pnp_copy_f32(D.ptr(ctx.z_out), D.ptr(out), out.numel(), D.stream()))

    def denoise(self, noisy, sigma_est=0):
        from ..engine import ProxCtx
        noisy = np.asarray(noisy)
        H, W = noisy.shape
        dev = D.require_cuda()
        z = D.to_lines(noisy, H, W, dev)
        out = torch.empty_like(z)
        self._dev_denoise(ProxCtx(z, out, H, W, sigma_est=float(sigma_est)))
        return D.from_lines(out, H, W).reshape(H, W)
