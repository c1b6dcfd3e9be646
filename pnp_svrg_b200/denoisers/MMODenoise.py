"""MMODenoiser -- same signature as the reference's denoisers/MMODenoise.py:105-128 (20-layer
conv3x3 + bias + LeakyReLU "DnCNN_nobn", output = net(x) + x, input and output clamped to [0, 1])."""
import sys

import numpy as np
import torch

from .. import device as D
from . import _cnn
from .denoiser import Denoise


def load_net(pth=None, net_type='DnCNN_nobn', channels=1, n_lev=0.01, cuda=True, root_folder='.'):
    """Unpickle the reference checkpoint (MMODenoise.py:42-71) and return the bare simple_CNN module."""
    from .models import basic_models
    if 'DnCNN_nobn' in net_type:
        pth = root_folder + 'checkpoints/pretrained/' + net_type + '_nch_' + str(channels) + '_nlev_' + str(n_lev) + '.pth'
    if pth is None:
        raise NameError('Could not load ' + str(net_type))
    saved = {k: sys.modules.get(k) for k in ('models', 'models.basic_models')}
    try:        # the pickle names the class as models.basic_models.simple_CNN
        pkg = type(sys)('models')
        pkg.basic_models = basic_models
        sys.modules['models'], sys.modules['models.basic_models'] = pkg, basic_models
        ckpt = torch.load(pth, map_location='cpu', weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return (ckpt.module if hasattr(ckpt, 'module') else ckpt).eval()


class MMODenoiser(Denoise):
    _uses_sigma_est = False

    def __init__(self, model=None, channels=3, path=None, cuda=True, sigma=0.01, root_path='.', *, precision='fp32'):
        super().__init__()
        self.sigma = sigma
        self.precision = {'fp32': 0, 'bf16': 1, 'bf16x3': 2}[precision]
        if model is None:
            if channels != 1:
                raise NotImplementedError('only the single-channel (grey) networks are built on the GPU path')
            model = load_net(path, net_type='DnCNN_nobn', channels=channels, cuda=cuda, root_folder=root_path,
                             n_lev=self.sigma)
        self.network = model.module if hasattr(model, 'module') else model
        dev = D.require_cuda()
        # the reference feeds np.moveaxis(noisy, -1, 0) of a 2-D image = its TRANSPOSE to the net
        # (MMODenoise.py:126-128); the device line layout already is that transpose -> no axis swap
        self.model = _cnn.PackedNet(_cnn.layers_from_simple_cnn(self.network), mode=1, swap_spatial=False, device=dev)

    def _dev_denoise(self, ctx):
        self.t += 1
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
