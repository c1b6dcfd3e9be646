"""Same exports as the reference's denoisers/__init__.py:4-8."""
from .denoiser import Denoise
from .NLM import NLMDenoiser
from .TV import TVDenoiser

__all__ = ['Denoise', 'NLMDenoiser', 'TVDenoiser']
