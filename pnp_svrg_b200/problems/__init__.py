"""Same exports as the reference's problems/__init__.py:4-7."""
from .problem import Problem, MiniBatch
from .CSMRI import CSMRI
from .DeblurSR import Deblur
from .PR import PhaseRetrieval

__all__ = ['Problem', 'CSMRI', 'Deblur', 'PhaseRetrieval']
