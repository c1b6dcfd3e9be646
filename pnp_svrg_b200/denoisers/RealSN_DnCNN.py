"""RealSN_DnCNNDenoiser -- same signature as the reference's denoisers/RealSN_DnCNN.py:8-40.  The
whole call (min/max normalisation, the 17-layer 3x3 conv stack with folded BatchNorm, the residual
subtraction and the un-normalisation) runs on the GPU without host round trips."""
import os

import numpy as np
import torch

from .. import device as D
from . import _cnn
from .denoiser import Denoise


def load_state_dict(model_type, sigma, weights_dir=None):
    """Same file the reference's load_model reads (DeepDenoisers/utils/utils.py:11, CWD-relative)."""
    name = model_type + "_noise" + str(sigma) + ".pth"
    cands = [os.path.join(weights_dir, name)] if weights_dir else []
    cands.append(os.path.join(".", "denoisers", "DeepDenoisers", "Pretrained_models", name))
    for p in cands:
        if os.path.exists(p):
            return torch.load(p, map_location='cpu')
    raise FileNotFoundError('pretrained weights %s not found (looked in %s)' % (name, cands))


class RealSN_DnCNNDenoiser(Denoise):
    _uses_sigma_est = False          # denoise() ignores sigma_est, as in the reference

    def __init__(self, model_type, sigma, *, state_dict=None, weights_dir=None, precision='fp32'):
        super().__init__()
        self.model_type = model_type
        self.sigma = sigma
        self.precision = {'fp32': 0, 'bf16': 1, 'bf16x3': 2}[precision]
        sd = state_dict if state_dict is not None else load_state_dict(model_type, sigma, weights_dir)
        dev = D.require_cuda()
        # denoisers/RealSN_DnCNN.py:27-29
        scale_range = 1.0 + self.sigma / 255.0 / 2.0
        scale_shift = (1 - scale_range) / 2.0
        self.model = _cnn.PackedNet(_cnn.layers_from_dncnn_state_dict(sd), mode=0, swap_spatial=True, device=dev,
                                    range_=scale_range, shift_in=scale_shift)

    def _dev_denoise(self, ctx):
        self.model.forward(ctx.z_in, ctx.z_out, ctx.W, ctx.H, xrec=ctx.xrec, mse_log=ctx.mse_log, slot=ctx.slot,
                           precision=self.precision)

    def denoise(self, noisy, sigma_est=0):
        from ..engine import ProxCtx
        noisy = np.asarray(noisy)
        m, n = noisy.shape
        dev = D.require_cuda()
        z = D.to_lines(noisy, m, n, dev)
        out = torch.empty_like(z)
        self._dev_denoise(ProxCtx(z, out, m, n))
        return D.from_lines(out, m, n).reshape(m, n)
