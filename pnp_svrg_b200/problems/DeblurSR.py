"""Deblur + super-resolution -- placeholder until the K2/K3 kernels land."""
from .problem import Problem


class Deblur(Problem):
    def __init__(self, img_path=None, H=64, W=64, kernel_path=None, kernel=None, scale_percent=50,
                 snr=None, sigma=None, *, image=None):
        raise NotImplementedError('Deblur: CUDA kernels not built yet in this revision (no CPU fallback)')
