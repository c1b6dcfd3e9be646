"""float64 NumPy restatement of the reference's inverse problems (data-fidelity gradients).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED: tests/test_oracle_vs_reference.py
and tests/golden/ref_*.npz compare every function here with the reference's own code
(problems/problem.py, CSMRI.py, DeblurSR.py, PR.py) executed under identical seeds.

Each class consumes NumPy's legacy global RNG in exactly the reference's call order
 so that ``np.random.seed(s); Port(...)`` and
``np.random.seed(s); Reference(...)`` build identical problems.
"""
import numpy as np

from . import pylops_port
from .skimage_port import peak_signal_noise_ratio


def load_image(img, H, W):
    """problems/problem.py:16-25 -- PIL open, resize((H, W)) (bicubic default), min-max to [0, 1].
    ``img`` may also be an already loaded 2-D array (used for the committed fixtures)."""
    if isinstance(img, str):
        from PIL import Image
        tmp = np.array(Image.open(img).resize((H, W)))
    elif img is None:
        raise Exception('Need to pass in image path or image')
    else:
        tmp = np.asarray(img)
        if tmp.shape != (H, W):
            from PIL import Image
            tmp = np.array(Image.fromarray(tmp).resize((H, W)))
    return (tmp - np.min(tmp)) / (np.max(tmp) - np.min(tmp))


class ProblemPort:
    def __init__(self, img, H, W):
        self.H, self.W, self.N = H, W, H * W
        self.M = self.N
        self.Xrec = load_image(img, H, W)
        self.X = self.Xrec.ravel()
        self.Xinit = np.empty_like(self.X)

    def PSNR(self, w):
        # problems/problem.py:33-35
        return np.around(peak_signal_noise_ratio(self.Xrec, np.asarray(w).reshape(self.H, self.W)),
                         decimals=2)

    def _noise_level(self):
        # problems/problem.py:37-61 (note: norm, not norm**2 -- kept as in the reference)
        nrm = np.linalg.norm(self.Y0.ravel())
        if self.snr is not None and self.sigma is None:
            self.sigma = np.sqrt(nrm / 10 ** (self.snr / 10) / self.H / self.W)
        elif self.sigma is not None and self.snr is None:
            if self.sigma > 0:
                self.snr = 10 * np.log10(nrm / self.sigma ** 2 / self.H / self.W)
            elif self.sigma == 0:
                self.snr = 10e9
            else:
                raise Exception('Sigma cannot be negative.')
        elif self.snr is None and self.sigma is None:
            self.sigma, self.snr = 0, 10e9
        else:
            raise Exception('Please specify either sigma (sigma) or signal-to-noise ratio (snr).')

    def select_mb(self, size):
        # problems/problem.py:110-117
        if size > self.M:
            print('MB size is too big: ', size, ' > ', self.M)
        batch = np.zeros(self.M)
        batch[np.random.choice(self.M, size, replace=False)] = 1
        return batch.astype(int)


class CSMRIPort(ProblemPort):
    """problems/CSMRI.py:12-89."""

    def __init__(self, img=None, H=256, W=256, sample_prob=0.5, snr=None, sigma=None):
        super().__init__(img, H, W)
        self.pname = 'csmri'
        self.sample_prob, self.snr, self.sigma = sample_prob, snr, sigma
        # :45 Bernoulli mask per k-space pixel
        self.mask = np.random.choice([0, 1], size=(H, W), p=[1 - sample_prob, sample_prob])
        # :47-59 the reference multiplies by a dense DFT matrix; fft2 equals it to ~6e-9
        # (SURVEY G2).  The dense product is kept so Y0 is bit-identical to the reference.
        i, j = np.meshgrid(np.arange(H), np.arange(W))
        F = np.power(np.exp(-2 * np.pi * 1J / H), i * j)
        self.Y0 = self.mask * (F.dot(self.Xrec).dot(F.T))
        self._noise_level()
        noises = np.random.normal(0, self.sigma, self.Y0.shape)
        self.Y = self.Y0 + self.mask * noises
        x0 = np.absolute(np.fft.ifft2(self.Y)).ravel()
        self.Xinit = (x0 - x0.min()) / (x0.max() - x0.min())
        self.lrH, self.lrW = H, W
        self.M = self.N
        self.M0 = np.count_nonzero(self.mask)

    def select_mb(self, size):
        # :66-74 draw from the sampled support only
        if size > self.M:
            print('MB size is too big: ', size, ' > ', self.M)
        batch = np.zeros(self.M)
        batch[np.random.choice(np.flatnonzero(self.mask), size, replace=False)] = 1
        return batch.reshape(self.H, self.W).astype(int)

    def _grad(self, z, sel):
        spec = np.fft.fft2(np.asarray(z).reshape(self.H, self.W)) * sel
        idx = np.nonzero(sel)
        spec[idx] = spec[idx] - self.Y[idx]
        return np.real(np.fft.ifft2(spec)).ravel()

    def grad_full(self, z):
        return self._grad(z, self.mask) / self.M0        # :76-81

    def grad_stoch(self, z, mb):
        return self._grad(z, self.mask * mb)             # :83-89 (caller divides by B)


class DeblurPort(ProblemPort):
    """problems/DeblurSR.py:17-147."""

    def __init__(self, img=None, H=64, W=64, kernel_path=None, kernel=None, scale_percent=50,
                 snr=None, sigma=None):
        super().__init__(img, H, W)
        self.pname = 'deblur'
        self.scale_percent, self.snr, self.sigma = scale_percent, snr, sigma
        if kernel_path is None and kernel is None:
            raise Exception('Need to pass in kernel path or kernel as image')
        N = self.N
        if kernel_path is not None:                      # :74-75
            from PIL import Image
            if isinstance(kernel_path, str):
                B = np.array(Image.open(kernel_path).resize((H, W)))
            else:
                B = np.array(Image.fromarray(np.asarray(kernel_path)).resize((H, W)))
        elif isinstance(kernel, str) and kernel == 'Identity':
            B = np.zeros(N)
            B[0] = 1
        elif isinstance(kernel, str) and kernel == 'Minimal':   # :80-87 (H used for both axes)
            B = np.zeros((H, W))
            B[0, 0] = 1
            B[H // 2, H // 2] = 1
            B[H // 2, H // 3] = 1
            B[H // 2, H // 4] = 1
            B /= 4
        else:
            B = np.asarray(kernel)
        self.B = B.ravel() / N                           # :93
        self.lrH = int(H * scale_percent / 100)
        self.lrW = int(W * scale_percent / 100)
        self.M = self.lrH * self.lrW
        eps = 1e-10
        if scale_percent == 100:                         # :97-108
            self.Bop = pylops_port.Identity(self.M)
        else:
            ptsH = np.linspace(eps, H - (1 + eps), self.lrH)
            ptsW = np.linspace(eps, W - (1 + eps), self.lrW)
            meshW, meshH = np.meshgrid(ptsH, ptsW)
            self.Bop = pylops_port.Bilinear(np.vstack([meshH.ravel(), meshW.ravel()]), (H, W))
        self.Y0 = self.forward_model(self.X)
        self._noise_level()
        self.Y = self.Y0 + np.random.normal(0, self.sigma, self.Y0.shape)
        self.Xinit = np.random.uniform(0.0, 1.0, N)
        self._Badj = np.roll(np.flip(self.B), 1)         # :132,147 adjoint (correlation) kernel

    def fft_blur(self, a, b):
        # :119-120 raveled length-N circular convolution, scaled by sqrt(N)
        return np.real(np.fft.ifft(np.fft.fft(a.ravel()) * np.fft.fft(b.ravel()))) * np.sqrt(self.N)

    def forward_model(self, w):
        return self.Bop * self.fft_blur(w, self.B)

    def grad_full(self, z):
        res = self.Bop * self.fft_blur(np.asarray(z).ravel(), self.B) - self.Y
        return self.fft_blur(self.Bop.H * res, self._Badj) / self.M

    def grad_stoch(self, z, mb):
        idx = np.nonzero(np.asarray(mb).ravel())
        down = self.Bop * self.fft_blur(np.asarray(z).ravel(), self.B)
        res = np.zeros(self.M)
        res[idx] = down[idx] - self.Y[idx]
        return self.fft_blur(self.Bop.H * res, self._Badj)


class PhaseRetrievalPort(ProblemPort):
    """problems/PR.py:13-87 (dense real Gaussian A, amplitude loss)."""

    def __init__(self, img=None, H=256, W=256, num_meas=-1, snr=None, sigma=None):
        super().__init__(img, H, W)
        self.pname = 'pr'
        self.M, self.snr, self.sigma = num_meas, snr, sigma
        self.A = np.random.randn(self.M, self.N)
        self.Y0 = np.absolute(self.A.dot(self.X)).ravel()
        self._noise_level()
        self.Y = self.Y0 + np.random.normal(0, self.sigma, self.Y0.shape)
        self._spec_init()
        self.Xinit = (self.Xinit - self.Xinit.min()) / (self.Xinit.max() - self.Xinit.min())

    def _spec_init(self):
        # :50-63 power iteration on D = A^T diag(Y) A / M, normalised by max (not norm)
        nrm = np.linalg.norm(self.X)
        D = self.A.T.dot(self.A * self.Y[:, None]) / self.M
        m, mold = 1, 2
        cur, old = 2 * np.ones(self.N), np.ones(self.N)
        while abs(m - mold) > 1e-5 and np.linalg.norm(cur - old) > 1e-5:
            mold, old = m, cur
            cur = D.dot(cur)
            m = np.max(cur)
            cur = cur / m
        self.Xinit = np.sqrt(m) * cur / np.linalg.norm(cur) * nrm

    def _grad(self, A, y, z):
        t = A.dot(np.asarray(z).ravel()).ravel()
        wgt = np.divide(np.absolute(t) - y, np.absolute(t))
        return A.T.dot(wgt * t).ravel()

    def grad_full(self, z):
        return self._grad(self.A, self.Y.ravel(), z) / self.M    # :75-79

    def grad_stoch(self, z, mb):
        idx = np.nonzero(mb)                                      # :81-87
        return self._grad(self.A[idx], self.Y[idx], z)


class CDPPort(ProblemPort):
    """Coded-diffraction phase retrieval, intensity loss: checker of the ADDITIVE mode
    PhaseRetrieval(model='cdp') (SURVEY section 8(a'); no counterpart in the reference, so PARITY UNPINNED --
    tests/test_oracle_golden.py checks the gradient against finite differences of f instead).
    Same RNG call order as the device class: codes, then the noise.

        A_l x = fft2(d_l o x)/sqrt(N), d_l = 1j**codes[l];  f(x) = sum((|Ax|^2 - y)^2) / (4M)
        grad_full(z) = Re(A^H((|Az|^2 - y) o Az)) / M ;  grad_stoch(z, mb) the same over the minibatch, no 1/B
    """

    def __init__(self, img=None, H=256, W=256, n_masks=4, snr=None, sigma=None):
        super().__init__(img, H, W)
        self.pname = 'pr'
        self.L, self.snr, self.sigma = n_masks, snr, sigma
        self.M = self.L * self.N
        self.codes = np.random.randint(0, 4, size=(self.L, H, W))
        self.d = 1j ** self.codes
        self.Y0 = (np.abs(self.A_op(self.X)) ** 2).ravel()
        self._noise_level()
        self.Y = self.Y0 + np.random.normal(0, self.sigma, self.Y0.shape)
        self._spec_init()
        self.Xinit = (self.Xinit - self.Xinit.min()) / (self.Xinit.max() - self.Xinit.min())

    def A_op(self, w):
        return np.fft.fft2(self.d * np.asarray(w, dtype=np.float64).reshape(1, self.H, self.W)) / np.sqrt(self.N)

    def AH_op(self, r):
        return (np.conj(self.d) * np.fft.ifft2(r) * np.sqrt(self.N)).sum(0)

    def _spec_init(self):
        nrm = np.linalg.norm(self.X)
        Y = self.Y.reshape(self.L, self.H, self.W)
        m, mold = 1, 2
        cur, old = 2 * np.ones(self.N), np.ones(self.N)
        it = 0
        while abs(m - mold) > 1e-5 and np.linalg.norm(cur - old) > 1e-5 and it < 500:
            mold, old = m, cur
            cur = np.real(self.AH_op(Y * self.A_op(cur))).ravel() / self.M
            m = np.max(cur)
            cur = cur / m
            it += 1
        self.Xinit = np.sqrt(abs(m)) * cur / np.linalg.norm(cur) * nrm

    def f(self, w):
        return np.sum((np.abs(self.A_op(w)).ravel() ** 2 - self.Y) ** 2) / 4 / self.M

    def _grad(self, z, mb=None):
        t = self.A_op(z)
        q = np.abs(t) ** 2 - self.Y.reshape(t.shape)
        if mb is not None:
            q = q * np.asarray(mb).reshape(t.shape)
        return np.real(self.AH_op(q * t)).ravel()

    def grad_full(self, z):
        return self._grad(z) / self.M

    def grad_stoch(self, z, mb):
        return self._grad(z, mb)
