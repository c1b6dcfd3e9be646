"""BM3DDenoiser -- signature of the reference's denoisers/BM3D.py:9-25.  BM3D is a closed third-party
binary (bm3d 3.0.9) that is out of the hot-path scope: this class is a
host passthrough that needs the ``bm3d`` package; it is not a GPU kernel and is not used by bench.py."""
import numpy as np
import torch

from .. import device as D
from .denoiser import Denoise


class BM3DDenoiser(Denoise):
    _fused_psnr = False

    def __init__(self, decay=1, denoise_strength=0, sigma_modifier=1):
        super().__init__()
        self.decay = decay
        self.denoise_strength = denoise_strength
        self.sigma_modifier = sigma_modifier

    def denoise(self, noisy, sigma_est=0):
        try:
            from bm3d import bm3d
        except ImportError as e:
            raise ImportError('BM3DDenoiser needs the third-party bm3d package (closed binary, not bundled)') from e
        self.t += 1
        if sigma_est > 0:
            return bm3d(noisy, self.sigma_modifier * sigma_est)
        return bm3d(noisy, self.denoise_strength * self.decay ** self.t)

    def _dev_denoise(self, ctx):
        # host passthrough: device -> host -> bm3d -> device (synchronises)
        s = float(ctx.sig_log[0].item()) / ctx.W if ctx.sig_log is not None else float(ctx.sigma_est)
        if ctx.slot is not None and ctx.sig_log is not None:
            s = float(ctx.sig_log[int(ctx.slot.item())].item()) / ctx.W
        z = D.from_lines(ctx.z_in, ctx.H, ctx.W).reshape(ctx.H, ctx.W)
        out = self.denoise(z, sigma_est=s if s == s else 0.0)
        ctx.z_out.copy_(D.to_lines(np.asarray(out), ctx.H, ctx.W, ctx.z_out.device))
