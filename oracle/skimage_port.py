"""float64 NumPy restatement of the scikit-image 0.18.2 / PyWavelets 1.1.1 calls on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: scikit-image and
PyWavelets are absent from the build image (requirements.txt:28,31 of the reference pin
PyWavelets 1.1.1 / scikit-image 0.18.2); the functions below restate their published
algorithms.  Call sites in the reference:

  peak_signal_noise_ratio   problems/problem.py:35
  estimate_sigma            algorithms/pnp_gd.py:49, pnp_sgd.py:50, pnp_svrg.py:71,
                            pnp_saga.py:64, pnp_sarah.py:47,89
  denoise_wavelet           denoisers/TV.py:24,26
  denoise_nl_means          denoisers/NLM.py:25,27

The reference always calls the restoration functions on a 2-D (H, W) float64 image with
``multichannel=True``, which makes scikit-image treat every COLUMN as a channel holding a
1-D signal of length H.  The restatements keep that behaviour.
"""
import numpy as np

# Daubechies-2 decomposition high-pass filter (pywt.Wavelet('db2').dec_hi)
DB2_DEC_HI = np.array([-0.48296291314469025, 0.836516303737469,
                       -0.22414386804185735, -0.12940952255092145])
# scipy.stats.norm.ppf(0.75)
GAUSS_Q75 = 0.6744897501960817
SQRT1_2 = 0.7071067811865476


# --------------------------------------------------------------------------- PSNR
def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    """skimage.metrics.peak_signal_noise_ratio for float images.

    With ``data_range=None`` and a float ``image_true`` scikit-image uses 1 when
    ``image_true.min() >= 0`` and 2 otherwise (dtype range of floats is (-1, 1)).
    """
    image_true = np.asarray(image_true, dtype=np.float64)
    image_test = np.asarray(image_test, dtype=np.float64)
    if data_range is None:
        data_range = 1.0 if image_true.min() >= 0 else 2.0
    err = np.mean((image_true - image_test) ** 2)
    with np.errstate(divide='ignore'):
        return 10.0 * np.log10((data_range ** 2) / err)


# ------------------------------------------------------------------ 1-D DWT helpers
def _symmetric_ext_index(idx, n):
    """Half-sample symmetric extension index (pywt mode 'symmetric'):
    ... x1 x0 | x0 x1 ... x[n-1] | x[n-1] x[n-2] ...  (period 2n)."""
    idx = np.mod(idx, 2 * n)
    return np.where(idx < n, idx, 2 * n - 1 - idx)


def dwt_detail_db2_axis0(x):
    """Single-level db2 detail coefficients along axis 0, mode 'symmetric'.

    pywt's downsampling convolution: out[o] = sum_j h[j] * x_ext[2*o + 1 - j],
    o = 0 .. (n + 3)//2 - 1 (len(h) = 4)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    n_out = (n + 3) // 2
    o = np.arange(n_out)
    out = np.zeros((n_out,) + x.shape[1:], dtype=np.float64)
    for j in range(4):
        src = _symmetric_ext_index(2 * o + 1 - j, n)
        out += DB2_DEC_HI[j] * x[src]
    return out


def sigma_mad_columns(z0):
    """Per-column sigma = median(|d| over d != 0) / Phi^-1(0.75)   (skimage _sigma_est_dwt)."""
    d = dwt_detail_db2_axis0(z0)
    a = np.abs(d)
    a[d == 0] = np.nan
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            med = np.nanmedian(a, axis=0)
    return med / GAUSS_Q75


def estimate_sigma(image, average_sigmas=False, multichannel=False):
    """skimage.restoration.estimate_sigma.

    multichannel=True on an (H, W) array: one estimate per column (db2 detail, MAD),
    averaged when ``average_sigmas``."""
    image = np.asarray(image, dtype=np.float64)
    if multichannel:
        sig = sigma_mad_columns(image.reshape(image.shape[0], -1) if image.ndim == 2 else image)
        if average_sigmas:
            return float(np.mean(sig))
        return list(sig)
    if image.ndim != 1:
        raise NotImplementedError('oracle restates only the 1-D (per-channel) estimate')
    return float(sigma_mad_columns(image[:, None])[0])


def _haar_fwd_axis0(x):
    n = x.shape[0]
    if n % 2:
        x = np.concatenate([x, x[-1:]], axis=0)      # symmetric extension of one sample
    e, o = x[0::2], x[1::2]
    return (e + o) * SQRT1_2, (e - o) * SQRT1_2


def _haar_inv_axis0(a, d):
    if a.shape[0] == d.shape[0] + 1:                 # pywt waverecn trims the approximation
        a = a[:-1]
    out = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=np.float64)
    out[0::2] = (a + d) * SQRT1_2
    out[1::2] = (a - d) * SQRT1_2
    return out


def haar_levels(n):
    """max(pywt.dwtn_max_level((n,), 'db1') - 3, 1) as in skimage _wavelet_threshold."""
    return max(int(np.floor(np.log2(n))) - 3, 1) if n >= 2 else 1


def bayes_shrink_columns(image, sigma):
    """Per-column multi-level Haar BayesShrink soft thresholding (skimage _wavelet_threshold
    applied to each image[..., c], which for a 2-D input is the 1-D column c).

    ``sigma`` is a scalar or one value per column."""
    x = np.asarray(image, dtype=np.float64)
    n = x.shape[0]
    sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), x.shape[1:])
    var = sigma ** 2
    eps = np.finfo(np.float64).eps
    levels = haar_levels(n)
    a = x
    details = []
    for _ in range(levels):
        a, d = _haar_fwd_axis0(a)
        details.append(d)
    for d in reversed(details):
        dvar = np.mean(d * d, axis=0)
        thr = var / np.sqrt(np.maximum(dvar - var, eps))
        mag = np.abs(d)
        with np.errstate(divide='ignore', invalid='ignore'):
            gain = 1.0 - thr / mag
        # pywt.threshold(mode='soft') would give NaN for 0/0 (thr == 0 and d == 0); a zero
        # coefficient stays zero in every other case, so it is kept at zero here.
        gain = np.where(mag > 0, np.maximum(gain, 0.0), 0.0)
        a = _haar_inv_axis0(a, d * gain)
    return a[:n]


def denoise_wavelet(image, sigma=None, wavelet='db1', mode='soft', wavelet_levels=None,
                    multichannel=False, convert2ycbcr=False, method='BayesShrink',
                    rescale_sigma=True):
    """skimage.restoration.denoise_wavelet, restricted to what denoisers/TV.py:24,26 uses."""
    if wavelet != 'db1' or mode != 'soft' or method != 'BayesShrink' or wavelet_levels is not None \
            or convert2ycbcr or not multichannel:
        raise NotImplementedError('oracle restates only the db1/soft/BayesShrink/multichannel call')
    image = np.asarray(image)
    if image.dtype.kind != 'f':
        raise NotImplementedError('float input only (the reference passes float64)')
    if sigma is None:
        raise NotImplementedError('sigma=None path is not on the reference hot path')
    return bayes_shrink_columns(image.astype(np.float64), sigma)


# ----------------------------------------------------------------------------- TV (Chambolle)
def denoise_tv_chambolle(image, weight=0.1, eps=2.e-4, n_iter_max=200):
    """skimage.restoration.denoise_tv_chambolle (0.18.2, `_denoise_tv_chambolle_nd`) for a 2-D float image.

    NOT on the reference's path (its TVDenoiser is the wavelet shrink above): checker of the additive
    `TVDenoiser(method='chambolle')` mode, SURVEY section 8(a').  PARITY UNPINNED (no skimage here);
    self-checks in tests/test_skimage_port.py: constant images are fixed points, the mean is preserved,
    total variation decreases.  `eps=0` never stops early, which is what the CUDA kernel does.
    """
    image = np.asarray(image, dtype=np.float64)
    ndim = image.ndim
    p = np.zeros((ndim,) + image.shape)
    g = np.zeros_like(p)
    d = np.zeros_like(image)
    out = image
    i = 0
    E_init = E_previous = 0.0
    while i < n_iter_max:
        if i > 0:
            d = -p.sum(0)                       # d is the (negative) divergence of p
            d[1:, :] += p[0, :-1, :]
            d[:, 1:] += p[1, :, :-1]
            out = image + d
        else:
            out = image
        E = (d ** 2).sum()
        g[0, :-1, :] = np.diff(out, axis=0)     # forward differences, last row / column stay 0
        g[1, :, :-1] = np.diff(out, axis=1)
        norm = np.sqrt((g ** 2).sum(axis=0))[np.newaxis, ...]
        E += weight * norm.sum()
        tau = 1.0 / (2.0 * ndim)
        norm *= tau / weight
        norm += 1.0
        p -= tau * g
        p /= norm
        E /= float(image.size)
        if i == 0:
            E_init = E
            E_previous = E
        else:
            if np.abs(E_previous - E) < eps * E_init:
                break
            E_previous = E
        i += 1
    return out


# ----------------------------------------------------------------------------- NLM
def denoise_nl_means(image, patch_size=7, patch_distance=11, h=0.1, multichannel=False,
                     fast_mode=True, sigma=0.0, exp_mode='exact'):
    """skimage.restoration.denoise_nl_means, classic (fast_mode=False) 2-D single-channel path.

    Restated from _nl_means_denoising_2d: even patch sizes are bumped to the next odd one,
    the image is reflect-padded by s//2, patch weights are a Gaussian of width (s-1)/4
    normalised by (sum(w) * h^2), the expected noise term 2*sigma^2 is subtracted from
    every squared difference, the accumulated distance is tested against the cut-off 5.0
    at the START of every patch row (early exit -> weight 0), and the weight is
    exp(-max(0, distance)).  The search window is clipped at the image border.

    ``exp_mode='exact'`` uses libm exp; scikit-image's own build may use a Schraudolph
    style ``fast_exp`` (not verifiable here) -- see DESIGN.md.
    """
    if fast_mode:
        raise NotImplementedError('the reference uses fast_mode=False (denoisers/NLM.py:11)')
    img = np.asarray(image, dtype=np.float64)
    if img.ndim != 2:
        raise NotImplementedError('2-D grey images only')
    s = patch_size + 1 if patch_size % 2 == 0 else patch_size
    d = patch_distance
    off = s // 2
    H, W = img.shape
    var = 2.0 * sigma * sigma
    A = (s - 1.0) / 4.0
    r = np.arange(-off, off + 1, dtype=np.float64)
    gr, gc = np.meshgrid(r, r, indexing='ij')
    w = np.exp(-(gr * gr + gc * gc) / (2 * A * A))
    w *= 1.0 / (np.sum(w) * h * h)
    pad = np.pad(img, off, mode='reflect')
    acc = np.zeros((H, W))
    wsum = np.zeros((H, W))
    rows = np.arange(H)[:, None]
    cols = np.arange(W)[None, :]
    for di in range(-d, d + 1):
        for dj in range(-d, d + 1):
            # candidate (row+di, col+dj) is inside the clipped window iff it is inside the image
            valid = (rows + di >= 0) & (rows + di < H) & (cols + dj >= 0) & (cols + dj < W)
            if not valid.any():
                continue
            # shifted copy of the padded image so that shifted[r, c] = pad[r+di, c+dj]
            sh = np.roll(pad, (-di, -dj), axis=(0, 1))
            diff2 = (pad - sh) ** 2 - var
            dist = np.zeros((H, W))
            dead = np.zeros((H, W), dtype=bool)
            for pi in range(s):
                dead |= dist > 5.0
                for pj in range(s):
                    dist = dist + w[pi, pj] * diff2[pi:pi + H, pj:pj + W]
            wt = np.exp(-np.maximum(0.0, dist))
            wt[dead | ~valid] = 0.0
            centre = sh[off:off + H, off:off + W]
            acc += wt * centre
            wsum += wt
    return acc / wsum
