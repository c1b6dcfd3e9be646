"""pnp_svrg_b200 -- B200-native (sm_100a) implementation of the PnP iteration hot path of
vmonardo/pnp-svrg behind the reference's own call signatures.

    from pnp_svrg_b200.problems import CSMRI, Deblur, PhaseRetrieval
    from pnp_svrg_b200.denoisers import TVDenoiser, NLMDenoiser, RealSN_DnCNNDenoiser
    from pnp_svrg_b200.algorithms import pnp_gd, pnp_sgd, pnp_svrg, pnp_saga, pnp_sarah

``install_as_reference()`` additionally registers the three sub-packages under the reference's
top-level names (``problems``, ``denoisers``, ``algorithms``) so existing scripts and notebooks
(``from problems import *`` ...) run unchanged.

The compute path is hand-written CUDA reached through the C ABI in include/pnp_b200.h; there is
no CPU fallback -- without the built library or without a GPU every compute call raises.
"""
import importlib
import sys

__version__ = '0.1.0'


def install_as_reference():
    """Alias pnp_svrg_b200.{problems,denoisers,algorithms} as top-level packages."""
    for name in ('problems', 'denoisers', 'algorithms'):
        mod = importlib.import_module('pnp_svrg_b200.' + name)
        existing = sys.modules.get(name)
        if existing is not None and existing is not mod:
            raise RuntimeError('a different top-level package %r is already imported' % name)
        sys.modules[name] = mod
        for sub, m in list(sys.modules.items()):
            if sub.startswith('pnp_svrg_b200.' + name + '.'):
                sys.modules[name + sub[len('pnp_svrg_b200.' + name):]] = m
