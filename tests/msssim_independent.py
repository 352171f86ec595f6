"""Independent fp64 MS-SSIM (test helper): written from the published definitions, NOT from the oracle's code, to cross-check
``oracle.eovae_oracle.ms_ssim`` (the reference delegates MS-SSIM to ``torchmetrics``, which does not exist in this image, so
no reference-held value can pin it; consistency_loss.py:24-37 is the call site).

* SSIM / contrast-structure maps: Wang, Bovik, Sheikh, Simoncelli, "Image quality assessment: from error visibility to
  structural similarity", IEEE TIP 2004, eqs. (6), (9), (10), (13): Gaussian-weighted local means / variances / covariance,
  C1 = (K1 L)^2, C2 = (K2 L)^2 with K1 = 0.01, K2 = 0.03, L = data_range.
* Multi-scale combination: Wang, Simoncelli, Bovik, "Multi-scale structural similarity for image quality assessment",
  Asilomar 2003, eq. (7): prod_{j<M} cs_j^beta_j * ssim_M^beta_M, betas (0.0448, 0.2856, 0.3001, 0.2363, 0.1333).
* torchmetrics' conventions at the reference's call (gaussian_kernel=True, sigma=1.5, reduction mean, normalize='relu'):
  window length 2*int(3.5*sigma + 0.5) + 1 = 11 (derived from sigma; the kernel_size=5 argument is not the window), only
  window positions that lie entirely inside the image contribute ("valid" filtering: its reflect padding is cropped away
  again), per-sample mean of the maps over (C, H, W), relu on every per-scale value, 2x2 average pooling between scales.

Structure here (separable 1-D correlations with scipy.ndimage on float64 numpy arrays, per image and channel) shares no code
with the oracle (one depthwise 2-D conv over a stacked fp32 torch tensor)."""
import numpy as np
from scipy.ndimage import correlate1d

BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _window(sigma=1.5):
    half = int(3.5 * sigma + 0.5)
    d = np.arange(-half, half + 1, dtype=np.float64)
    w = np.exp(-0.5 * (d / sigma) ** 2)
    return w / w.sum(), half


def _local_mean(a, w, half):
    """Gaussian-weighted mean over the 2-D window, only where the window fits inside the image."""
    f = correlate1d(correlate1d(a, w, axis=-2, mode="constant"), w, axis=-1, mode="constant")
    return f[..., half:-half, half:-half]


def ssim_cs_per_sample(x, y, data_range):
    """x, y: float64 [B, C, H, W] -> (mean SSIM, mean contrast-structure) per sample."""
    w, half = _window()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mx, my = _local_mean(x, w, half), _local_mean(y, w, half)
    vx = np.maximum(_local_mean(x * x, w, half) - mx * mx, 0.0)
    vy = np.maximum(_local_mean(y * y, w, half) - my * my, 0.0)
    cxy = _local_mean(x * y, w, half) - mx * my
    cs = (2.0 * cxy + c2) / (vx + vy + c2)
    lum = (2.0 * mx * my + c1) / (mx * mx + my * my + c1)
    b = x.shape[0]
    return (lum * cs).reshape(b, -1).mean(1), cs.reshape(b, -1).mean(1)


def _pool2(a):
    b, c, h, w = a.shape
    return a[..., : h // 2 * 2, : w // 2 * 2].reshape(b, c, h // 2, 2, w // 2, 2).mean(axis=(3, 5))


def ms_ssim(x, y, data_range=6.0, betas=BETAS):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    out = np.ones(x.shape[0])
    for j, beta in enumerate(betas):
        s, cs = ssim_cs_per_sample(x, y, data_range)
        v = s if j == len(betas) - 1 else cs
        out = out * np.maximum(v, 0.0) ** beta
        x, y = _pool2(x), _pool2(y)
    return float(out.mean())
