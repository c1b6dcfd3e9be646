"""pnp_svrg_b200 -- B200-native (sm_100a) implementation of the PnP iteration hot path of
vmonardo/pnp-svrg behind the reference's own call signatures.

    from pnp_svrg_b200.problems import CSMRI, Deblur, PhaseRetrieval
    from pnp_svrg_b200.denoisers import TVDenoiser, NLMDenoiser, RealSN_DnCNNDenoiser
    from pnp_svrg_b200.algorithms import pnp_gd, pnp_sgd, pnp_svrg, pnp_saga, pnp_sarah

``install_as_reference()`` additionally registers the three sub-packages under the reference's
top-level names (``problems``, ``denoisers``, ``algorithms``) and ``Utilities`` (``display_results``) so
existing scripts and notebooks (``from problems import *``, ``from Utilities import display_results`` ...)
run unchanged; with ``hyperopt=True`` and no hyperopt installed, ``pnp_svrg_b200.search`` also answers
``from hyperopt import fmin, tpe, hp, Trials`` (the subset the sweep scripts use).

The compute path is hand-written CUDA reached through the C ABI in include/pnp_b200.h; there is
no CPU fallback -- without the built library or without a GPU every compute call raises.
"""
import importlib
import sys

__version__ = '0.1.0'


def install_as_reference(hyperopt=False):
    """Alias pnp_svrg_b200.{problems,denoisers,algorithms,Utilities} as top-level modules."""
    for name in ('problems', 'denoisers', 'algorithms', 'Utilities'):
        mod = importlib.import_module('pnp_svrg_b200.' + name)
        existing = sys.modules.get(name)
        if existing is not None and existing is not mod:
            raise RuntimeError('a different top-level package %r is already imported' % name)
        sys.modules[name] = mod
        for sub, m in list(sys.modules.items()):
            if sub.startswith('pnp_svrg_b200.' + name + '.'):
                sys.modules[name + sub[len('pnp_svrg_b200.' + name):]] = m
    if hyperopt:
        try:
            importlib.import_module('hyperopt')
        except ImportError:
            import types
            search = importlib.import_module('pnp_svrg_b200.search')
            top = types.ModuleType('hyperopt')
            for k in ('fmin', 'tpe', 'rand', 'hp', 'Trials', 'space_eval', 'STATUS_OK', 'STATUS_FAIL'):
                setattr(top, k, getattr(search, k))
            hp_mod = types.ModuleType('hyperopt.hp')                 # from hyperopt.hp import quniform
            for k in ('uniform', 'quniform', 'loguniform', 'randint', 'choice'):
                setattr(hp_mod, k, getattr(search.hp, k))
            pyll = types.ModuleType('hyperopt.pyll')                  # from hyperopt.pyll import scope
            pyll.scope = search.scope
            sys.modules.update({'hyperopt': top, 'hyperopt.hp': hp_mod, 'hyperopt.pyll': pyll})
