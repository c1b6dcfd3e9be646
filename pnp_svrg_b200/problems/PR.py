"""Phase retrieval -- placeholder until the K4 kernels land."""
from .problem import Problem


class PhaseRetrieval(Problem):
    def __init__(self, img_path=None, H=256, W=256, num_meas=-1, snr=None, sigma=None, *, image=None):
        raise NotImplementedError('PhaseRetrieval: CUDA kernels not built yet in this revision (no CPU fallback)')
