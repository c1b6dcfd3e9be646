"""float64 NumPy restatement of the two pylops 1.14.0 operators the reference uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: pylops is absent
from the build image (reference requirements.txt:25 pins 1.14.0).  Call sites:
problems/DeblurSR.py:98 (Identity), :108 (signalprocessing.Bilinear), and the ``Bop * x`` /
``Bop.H * r`` products at :112,129,131,143,146.
"""
import numpy as np


class _Adjoint:
    def __init__(self, op):
        self._op = op

    def __mul__(self, x):
        return self._op.rmatvec(np.asarray(x))

    matvec = __mul__


class _LinOp:
    @property
    def H(self):
        return _Adjoint(self)

    def __mul__(self, x):
        return self.matvec(np.asarray(x))


class Identity(_LinOp):
    """pylops.Identity(N): y = x."""

    def __init__(self, N, M=None, dtype='float64', inplace=True):
        self.shape = (N if M is None else N, N if M is None else M)

    def matvec(self, x):
        return x.ravel()

    def rmatvec(self, y):
        return y.ravel()


class Bilinear(_LinOp):
    """pylops.signalprocessing.Bilinear(iava, dims).

    iava: (2, n) float coordinates (row, col) inside dims=(H, W).  Forward samples the
    raveled image with 4-tap bilinear weights; the adjoint scatter-adds the same weights.
    """

    def __init__(self, iava, dims, dtype='float64'):
        iava = np.asarray(iava, dtype=np.float64)
        if np.unique(iava, axis=1).shape[1] != iava.shape[1]:
            raise ValueError('repeated values in iava array')
        self.dims = tuple(dims)
        self.n = iava.shape[1]
        self.t = np.floor(iava[0]).astype(np.int64)
        self.l = np.floor(iava[1]).astype(np.int64)
        self.wr = iava[0] - self.t            # weight of the lower row (t + 1)
        self.wc = iava[1] - self.l            # weight of the right column (l + 1)
        self.shape = (self.n, int(np.prod(dims)))

    def matvec(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(self.dims)
        t, l, wr, wc = self.t, self.l, self.wr, self.wc
        return (x[t, l] * (1 - wr) * (1 - wc) + x[t, l + 1] * (1 - wr) * wc
                + x[t + 1, l] * wr * (1 - wc) + x[t + 1, l + 1] * wr * wc)

    def rmatvec(self, y):
        y = np.asarray(y, dtype=np.float64).ravel()
        t, l, wr, wc = self.t, self.l, self.wr, self.wc
        out = np.zeros(self.dims, dtype=np.float64)
        np.add.at(out, (t, l), y * (1 - wr) * (1 - wc))
        np.add.at(out, (t, l + 1), y * (1 - wr) * wc)
        np.add.at(out, (t + 1, l), y * wr * (1 - wc))
        np.add.at(out, (t + 1, l + 1), y * wr * wc)
        return out.ravel()
