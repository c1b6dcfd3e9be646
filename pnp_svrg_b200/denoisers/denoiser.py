"""Denoise base class (reference denoisers/denoiser.py:2-7) plus the device protocol the
iteration engine uses: ``_dev_denoise(ctx)`` works in place on GPU memory."""


class Denoise():
    # does denoise() read the sigma_est the algorithm hands it?  (lets the engine skip the
    # estimate_sigma kernel for denoisers that ignore it, e.g. the CNN wrappers)
    _uses_sigma_est = True

    def __init__(self):
        self.t = 0

    def denoise(self, noisy):
        raise NotImplementedError('Need to implement denoise() method')

    def _dev_denoise(self, ctx):
        """ctx: engine.ProxCtx -- z_in/z_out device line tensors, H, W, sigma source, PSNR sink."""
        raise NotImplementedError('%s has no device implementation (no CPU fallback)' % type(self).__name__)
