"""Same exports as the reference's denoisers/__init__.py:4-8 (MMODenoiser is imported explicitly by
the reference's drivers: ``from denoisers.MMODenoise import MMODenoiser``, pnp_csmri.py:7)."""
from .denoiser import Denoise
from .BM3D import BM3DDenoiser
from .RealSN_DnCNN import RealSN_DnCNNDenoiser
from .NLM import NLMDenoiser
from .TV import TVDenoiser

__all__ = ['Denoise', 'BM3DDenoiser', 'RealSN_DnCNNDenoiser', 'NLMDenoiser', 'TVDenoiser']
