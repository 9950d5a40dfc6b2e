"""Shared helpers for the parity tests (seeded synthetic inputs, comparison)."""
import numpy as np


def rng(seed=1701):     # seed mirrors the reference's GradientChecker (test_gradient_check_util.hpp:25)
    return np.random.default_rng(seed)


def smooth_images(r, n, h, w, shift=(1.7, -2.3)):
    """Two BGR frames with integer values 0..255 as scripts/run-flownet.py:30-35 feeds them: low-pass noise,
    second frame = first shifted by a sub-pixel offset plus noise."""
    import scipy.ndimage as ndi
    base = r.standard_normal((n, 3, h + 16, w + 16))
    base = ndi.gaussian_filter(base, sigma=(0, 0, 3, 3))
    base = (base - base.min()) / (base.max() - base.min())
    a = base[:, :, 8:8 + h, 8:8 + w]
    b = ndi.shift(base, (0, 0, shift[1], shift[0]), order=1, mode="nearest")[:, :, 8:8 + h, 8:8 + w]
    b = b + 0.01 * r.standard_normal(b.shape)
    to8 = lambda x: np.clip(np.round(x * 255), 0, 255).astype(np.float32)
    return to8(a), to8(b)


def maxabs(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert (nan_a == nan_b).all(), "NaN pattern differs"
    d = np.abs(np.where(nan_a, 0, a) - np.where(nan_b, 0, b))
    return float(d.max()) if d.size else 0.0
