"""Same exports as the reference's algorithms/__init__.py:4-8, plus ``SvrgRun`` (the device-resident state behind
``pnp_svrg``: ``SvrgRun(...).epoch()`` enqueues one whole SVRG epoch as one CUDA graph)."""
from .pnp_gd import pnp_gd, tune_pnp_gd
from .pnp_sgd import pnp_sgd, tune_pnp_sgd
from .pnp_svrg import pnp_svrg, tune_pnp_svrg
from ._loops import SvrgRun
from .pnp_saga import pnp_saga, tune_pnp_saga
from .pnp_sarah import pnp_sarah, tune_pnp_sarah

__all__ = ['pnp_gd', 'tune_pnp_gd', 'pnp_sgd', 'tune_pnp_sgd', 'pnp_svrg', 'tune_pnp_svrg',
           'pnp_saga', 'tune_pnp_saga', 'pnp_sarah', 'tune_pnp_sarah', 'SvrgRun']
