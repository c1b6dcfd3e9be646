"""Import the UNMODIFIED reference from /root/reference inside the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by oracle/gen_golden.py and by
tests that are skipped when /root/reference is absent (it does not exist on the GPU box).

The reference needs four things the image lacks:
  1. ``numpy.lib.npyio.save``  (problems/problem.py:5; gone in NumPy 2)
  2. ``skimage.metrics`` / ``skimage.restoration``   -> oracle.skimage_port
  3. ``pylops`` (``Identity``, ``signalprocessing.Bilinear``)  -> oracle.pylops_port
  4. ``bm3d`` and ``hyperopt`` (import-time only; ``STATUS_OK`` constant)
Nothing in the reference tree is modified or copied; the stubs live only in sys.modules.
"""
import importlib
import os
import sys
import types

from . import pylops_port, skimage_port

REF_ROOT = '/root/reference'


def reference_available(root=REF_ROOT):
    return os.path.isdir(os.path.join(root, 'algorithms'))


def _install_stubs():
    import numpy as np
    import numpy.lib.npyio as npyio
    if not hasattr(npyio, 'save'):
        npyio.save = np.save

    sk = types.ModuleType('skimage')
    sk.__path__ = []
    met = types.ModuleType('skimage.metrics')
    met.peak_signal_noise_ratio = skimage_port.peak_signal_noise_ratio
    res = types.ModuleType('skimage.restoration')
    res.estimate_sigma = skimage_port.estimate_sigma
    res.denoise_wavelet = skimage_port.denoise_wavelet
    res.denoise_nl_means = skimage_port.denoise_nl_means
    sk.metrics, sk.restoration = met, res
    sys.modules.setdefault('skimage', sk)
    sys.modules.setdefault('skimage.metrics', met)
    sys.modules.setdefault('skimage.restoration', res)

    pl = types.ModuleType('pylops')
    pl.__path__ = []
    sp = types.ModuleType('pylops.signalprocessing')
    sp.Bilinear = pylops_port.Bilinear
    pl.Identity = pylops_port.Identity
    pl.signalprocessing = sp
    sys.modules.setdefault('pylops', pl)
    sys.modules.setdefault('pylops.signalprocessing', sp)

    b3 = types.ModuleType('bm3d')

    def _no_bm3d(*a, **k):
        raise RuntimeError('bm3d is a closed binary that is not installed (out of scope)')
    b3.bm3d = _no_bm3d
    sys.modules.setdefault('bm3d', b3)

    ho = types.ModuleType('hyperopt')
    ho.STATUS_OK = 'ok'
    sys.modules.setdefault('hyperopt', ho)


class FakeClock:
    """Stands in for the ``time`` module inside algorithms/pnp_*.py.

    The reference loops are wall-clock bounded (``while time.time() - elapsed < tt``,
    e.g. algorithms/pnp_svrg.py:26,42).  ``tick()`` is called once per denoiser call, so
    ``tt = K`` gives exactly K denoiser calls."""

    def __init__(self):
        self.now = 0.0

    def time(self):
        return self.now

    def tick(self, dt=1.0):
        self.now += dt


class Reference:
    """Handle on the imported reference packages plus a trajectory recorder."""

    def __init__(self, root=REF_ROOT):
        if not reference_available(root):
            raise FileNotFoundError(root)
        _install_stubs()
        for name in ('problems', 'algorithms', 'denoisers'):
            if name in sys.modules and not getattr(sys.modules[name], '__file__', '').startswith(root):
                raise RuntimeError('a non-reference top-level package %r is already imported' % name)
        if root not in sys.path:
            sys.path.insert(0, root)
        self.problems = importlib.import_module('problems')
        self.algorithms = importlib.import_module('algorithms')
        self.clock = FakeClock()
        for mod in ('pnp_gd', 'pnp_sgd', 'pnp_svrg', 'pnp_saga', 'pnp_sarah'):
            importlib.import_module('algorithms.' + mod).time = self.clock
        # denoisers/__init__.py imports torch-based wrappers; import TV/NLM directly so that a
        # torch import problem cannot break the oracle
        self.TV = importlib.import_module('denoisers.TV')
        self.NLM = importlib.import_module('denoisers.NLM')

    def record(self, problem, denoiser):
        """Wrap select_mb / denoise so a run leaves its minibatches and iterates behind."""
        import numpy as np
        log = {'mb': [], 'noisy': [], 'denoised': [], 'sigma_est': []}
        sel, den = problem.select_mb, denoiser.denoise
        clock = self.clock

        def select_mb(size):
            mb = sel(size)
            log['mb'].append(np.flatnonzero(np.asarray(mb).ravel()).astype(np.int32))
            return mb

        def denoise(noisy, sigma_est=0):
            out = den(noisy=noisy, sigma_est=sigma_est)
            log['noisy'].append(np.array(noisy, dtype=np.float64))
            log['denoised'].append(np.array(out, dtype=np.float64))
            log['sigma_est'].append(float(sigma_est))
            clock.tick()
            return out

        problem.select_mb = select_mb
        denoiser.denoise = denoise
        return log

    def run(self, algo, problem, denoiser, budget, **kw):
        """Run algorithms.<algo> for exactly ``budget`` denoiser calls."""
        self.clock.now = 0.0
        fn = getattr(self.algorithms, algo)
        return fn(problem, denoiser, tt=float(budget), verbose=False, **kw)
