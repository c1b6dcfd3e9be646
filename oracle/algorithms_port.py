"""float64 NumPy restatement of the five PnP loops with a deterministic iteration budget.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against algorithms/pnp_*.py of
the reference run under a fake clock (oracle/refshim.py): same minibatch draws, same
iterates, same PSNR log (tests/test_oracle_vs_reference.py, tests/golden/ref_*.npz).

The reference bounds its loops by wall clock (``while time.time() - elapsed < tt``).  Here
``budget`` counts denoiser calls, which is what the fake clock turns ``tt`` into.

``vr_mode`` (pnp_svrg only):
  'as_committed'  v = mu                               algorithms/pnp_svrg.py:54
  'paper'         v = (g_B(z) - g_B(w)) / B + mu       algorithms/pnp_svrg.py:53 (commented out)
"""
import numpy as np

from .skimage_port import estimate_sigma

TOL = 1e-5


class DenoiserPort:
    def __init__(self):
        self.t = 0


class TVPort(DenoiserPort):
    """denoisers/TV.py:9-26 (wavelet BayesShrink, despite the name)."""

    def __init__(self, decay=1, denoise_strength=0, sigma_modifier=1):
        super().__init__()
        self.decay, self.denoise_strength, self.sigma_modifier = decay, denoise_strength, sigma_modifier

    def denoise(self, noisy, sigma_est=0):
        from .skimage_port import bayes_shrink_columns
        self.t += 1
        s = sigma_est * self.sigma_modifier if sigma_est > 0 else self.denoise_strength * self.decay ** self.t
        return bayes_shrink_columns(noisy, s)


class ChambollePort(DenoiserPort):
    """Additive mode TVDenoiser(method='chambolle') (SURVEY section 8(a')); weight as in the device class."""

    def __init__(self, weight=None, n_iter=20, decay=1, denoise_strength=0, sigma_modifier=1):
        super().__init__()
        self.weight, self.n_iter = weight, n_iter
        self.decay, self.denoise_strength, self.sigma_modifier = decay, denoise_strength, sigma_modifier

    def denoise(self, noisy, sigma_est=0):
        from .skimage_port import denoise_tv_chambolle
        self.t += 1
        if self.weight is not None:
            w = self.weight
        else:
            w = sigma_est * self.sigma_modifier if sigma_est > 0 else self.denoise_strength * self.decay ** self.t
        if not w > 0:
            return np.array(noisy)
        return denoise_tv_chambolle(noisy, weight=w, eps=0.0, n_iter_max=self.n_iter)


class NLMPort(DenoiserPort):
    """denoisers/NLM.py:9-27 with ``self.sigma`` (never set in the reference -> AttributeError)
    replaced by the ``sigma_est > 0`` test the sibling denoisers use."""

    def __init__(self, decay=1, denoise_strength=0, patch_size=4, patch_distance=5, sigma_modifier=1):
        super().__init__()
        self.decay, self.denoise_strength, self.sigma_modifier = decay, denoise_strength, sigma_modifier
        self.patch_size, self.patch_distance = patch_size, patch_distance

    def denoise(self, noisy, sigma_est=0):
        from .skimage_port import denoise_nl_means
        self.t += 1
        if sigma_est > 0:
            s = sigma_est * self.sigma_modifier
            return denoise_nl_means(noisy, h=s, sigma=s, fast_mode=False,
                                    patch_size=self.patch_size, patch_distance=self.patch_distance)
        return denoise_nl_means(noisy, h=self.denoise_strength * self.decay ** self.t, fast_mode=False,
                                patch_size=self.patch_size, patch_distance=self.patch_distance)


class IdentityPort(DenoiserPort):
    def denoise(self, noisy, sigma_est=0):
        self.t += 1
        return np.array(noisy)


class _Run:
    """Book-keeping shared by the five loops: budget clock, logs, prox step, stop rules."""

    def __init__(self, problem, denoiser, budget, converge_check, diverge_check, trace):
        self.p, self.d, self.budget = problem, denoiser, budget
        self.cc, self.dc = converge_check, diverge_check
        self.calls = 0
        self.psnr = []
        self.trace = trace

    def alive(self):
        return self.calls < self.budget

    def prox(self, z):
        z0 = np.copy(z).reshape(self.p.H, self.p.W)
        s = estimate_sigma(z0, multichannel=True, average_sigmas=True)
        out = self.d.denoise(noisy=z0, sigma_est=s)
        self.calls += 1
        if self.trace is not None:
            self.trace.append(np.array(out, dtype=np.float64).ravel())
        return out

    def stop(self, start_psnr):
        last = self.psnr[-1]
        if self.cc is True and np.abs(start_psnr - last) < TOL:
            return True
        if self.dc is True and last < 0:
            return True
        return False

    def result(self, z, name):
        return {'z': z, 'psnr_per_iter': self.psnr, 'algo_name': name, 'n_prox': self.calls}


def pnp_gd(problem, denoiser, eta, budget, lr_decay=1, converge_check=True, diverge_check=False,
           trace=None):
    r = _Run(problem, denoiser, budget, converge_check, diverge_check, trace)
    z = np.copy(problem.Xinit)
    i = 0
    r.psnr.append(problem.PSNR(z))
    while r.alive():
        start = problem.PSNR(z)
        z -= (eta * lr_decay ** i) * problem.grad_full(z)
        z0 = r.prox(z)
        r.psnr.append(problem.PSNR(z0))
        z = np.copy(z0).ravel()
        i += 1
        if r.stop(start):
            break
    return r.result(z, 'PnP GD')


def pnp_sgd(problem, denoiser, eta, budget, mini_batch_size, lr_decay=1, converge_check=True,
            diverge_check=False, trace=None):
    r = _Run(problem, denoiser, budget, converge_check, diverge_check, trace)
    z = np.copy(problem.Xinit)
    i = 0
    r.psnr.append(problem.PSNR(z))
    while r.alive():
        start = problem.PSNR(z)
        mb = problem.select_mb(mini_batch_size)
        z -= (eta * lr_decay ** i) * (problem.grad_stoch(z, mb) / mini_batch_size)
        z0 = r.prox(z)
        r.psnr.append(problem.PSNR(z0))
        z = np.copy(z0).ravel()
        i += 1
        if r.stop(start):
            break
    return r.result(z, 'PnP SGD')


def pnp_svrg(problem, denoiser, eta, budget, T2, mini_batch_size, lr_decay=1, converge_check=True,
             diverge_check=False, vr_mode='as_committed', trace=None):
    r = _Run(problem, denoiser, budget, converge_check, diverge_check, trace)
    z = np.copy(problem.Xinit)
    i = 0
    r.psnr.append(problem.PSNR(z))
    done = False
    while r.alive() and not done:
        mu = problem.grad_full(z)
        w = np.copy(z)
        r.psnr.append(problem.PSNR(z))
        for _ in range(T2):
            if not r.alive():
                break
            start = problem.PSNR(z)
            mb = problem.select_mb(mini_batch_size)       # drawn even when unused (pnp_svrg.py:52)
            if vr_mode == 'paper':
                v = (problem.grad_stoch(z, mb) - problem.grad_stoch(w, mb)) / mini_batch_size + mu
            elif vr_mode == 'as_committed':
                v = mu
            else:
                raise ValueError(vr_mode)
            z -= (eta * lr_decay ** i) * v
            z0 = r.prox(z)
            r.psnr.append(problem.PSNR(z0))
            z = np.copy(z0).ravel()
            if r.stop(start):
                done = True
                break
        i += 1
    return r.result(z, 'PnP SVRG')


def pnp_saga(problem, denoiser, eta, budget, mini_batch_size, hist_size=50, lr_decay=1,
             converge_check=True, diverge_check=False, trace=None):
    r = _Run(problem, denoiser, budget, converge_check, diverge_check, trace)
    z = np.copy(problem.Xinit)
    i = 0
    mb = problem.select_mb(mini_batch_size)
    g0 = problem.grad_stoch(z, mb) / mini_batch_size
    table = [g0] * hist_size
    prev = g0
    r.psnr.append(problem.PSNR(z))
    while r.alive():
        start = problem.PSNR(z)
        mb = problem.select_mb(mini_batch_size)
        slot = np.random.choice(hist_size, 1).item()
        table[slot] = problem.grad_stoch(z, mb) / mini_batch_size
        # non-textbook: subtracts the previous iteration's gradient (pnp_saga.py:47,72)
        v = table[slot].ravel() - prev.ravel() + sum(table).ravel() / hist_size
        z -= (eta * lr_decay ** i) * v
        z0 = r.prox(z)
        prev = table[slot]
        r.psnr.append(problem.PSNR(z0))
        z = np.copy(z0).ravel()
        i += 1
        if r.stop(start):
            break
    return r.result(z, 'pnp_saga')


def pnp_sarah(problem, denoiser, eta, budget, T2, mini_batch_size, lr_decay=1, converge_check=True,
              diverge_check=False, trace=None):
    r = _Run(problem, denoiser, budget, converge_check, diverge_check, trace)
    z = np.copy(problem.Xinit)
    i = 0
    done = False
    while r.alive() and not done:
        w_prev = np.copy(z)
        v_prev = problem.grad_full(z)
        w_next = r.prox(w_prev - eta * v_prev)            # no lr_decay here (pnp_sarah.py:36)
        r.psnr.append(problem.PSNR(w_next))
        w_next = w_next.ravel()
        for _ in range(T2):
            if not r.alive():
                break
            start = problem.PSNR(z)
            mb = problem.select_mb(mini_batch_size)
            v_next = (problem.grad_stoch(w_next, mb).ravel()
                      - problem.grad_stoch(w_prev, mb).ravel()) / mini_batch_size + v_prev.ravel()
            z -= (eta * lr_decay ** i) * v_next
            z0 = r.prox(z)
            v_prev = np.copy(v_next)
            w_prev = np.copy(z0).ravel()
            r.psnr.append(problem.PSNR(z0))
            z = np.copy(z0).ravel()
            if r.stop(start):
                done = True
                break
        i += 1
    return r.result(z, 'pnp_sarah')


class DnCNNPort(DenoiserPort):
    """denoisers/RealSN_DnCNN.py:8-40 with the network evaluated by torch on the CPU in fp32 (the
    reference moves it to CUDA; the arithmetic is the same).  ``layers``: list of (weight (co,ci,3,3),
    bn or None) with bn = (gamma, beta, mean, var); eps 1e-5, ReLU between layers, no bias."""

    def __init__(self, layers, sigma):
        super().__init__()
        self.layers, self.sigma = layers, sigma

    def _net(self, x):
        import torch
        import torch.nn.functional as F
        t = torch.from_numpy(x.astype(np.float32))[None, None]
        with torch.no_grad():
            for i, (w, bn) in enumerate(self.layers):
                t = F.conv2d(t, torch.from_numpy(np.asarray(w, dtype=np.float32)), padding=1)
                if bn is not None:
                    g, b, m, v = (torch.from_numpy(np.asarray(a, dtype=np.float32)) for a in bn)
                    t = F.batch_norm(t, m, v, g, b, training=False, eps=1e-5)
                if i < len(self.layers) - 1:
                    t = F.relu(t)
        return t[0, 0].numpy().astype(np.float64)

    def denoise(self, noisy, sigma_est=0):
        xt = np.copy(noisy)
        mn, mx = np.min(xt), np.max(xt)
        xt = (xt - mn) / (mx - mn)
        rng_ = 1.0 + self.sigma / 255.0 / 2.0
        sh = (1 - rng_) / 2.0
        xt = xt * rng_ + sh
        x = xt - self._net(xt)
        x = (x - sh) / rng_
        return x * (mx - mn) + mn
