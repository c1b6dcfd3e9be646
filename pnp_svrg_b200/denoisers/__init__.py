"""Same exports as the reference's denoisers/__init__.py:4-8."""
from .denoiser import Denoise
from .TV import TVDenoiser

__all__ = ['Denoise', 'TVDenoiser']
