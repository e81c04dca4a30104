"""CPU oracle for the quality metrics (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's cpu_baseline may import it).

Restates utils/metrics.py of the reference; that module cannot be imported here (it imports skimage / PIL at the top and
skimage is not installed), so:
  * ``fid_statistics`` / ``fid`` are the reference's own numpy / scipy calls (utils/metrics.py:83-87 and :102-116) -- the
    arithmetic lives in numpy 2.3 / scipy 1.18 (installed), so these ARE the reference's results for the same inputs;
    pinned by tests/golden/metrics.npz, which oracle/make_golden_metrics.py generates by running the reference's own
    MetricsCalculator.compute_fid_statistics / compute_fid (file loaded unmodified, skimage import stubbed);
  * ``ssim`` restates skimage.metrics.structural_similarity (scikit-image==0.21.0, requirements.txt:10; scipy==1.11.3 is
    pinned at :8) as called at utils/metrics.py:52 -- ``channel_axis=2, data_range=255`` and otherwise defaults:
    win_size 7, uniform filter (scipy.ndimage.uniform_filter, which skimage itself calls), K1 = 0.01, K2 = 0.03,
    use_sample_covariance=True, float64 arithmetic, mean over the image cropped by (win_size - 1) // 2, channel mean.
    PARITY UNPINNED against skimage itself (absent); pinned to the reference's own tests (tests/test_utils.py:30-41:
    identical images > 0.99, all-0 vs all-255 < 0.5) and to the closed form for constant images.
"""
import numpy as np
from scipy import linalg
from scipy.ndimage import uniform_filter


def fid_statistics(images):
    """utils/metrics.py:83-87.  images: array [N, ...] -> (mu, sigma)."""
    flat = np.asarray(images).reshape(len(images), -1)
    return np.mean(flat, axis=0), np.cov(flat, rowvar=False)


def fid(real_images, generated_images):
    """utils/metrics.py:100-116."""
    mu1, sigma1 = fid_statistics(real_images)
    mu2, sigma2 = fid_statistics(generated_images)
    diff = mu1 - mu2
    try:
        covmean, _ = linalg.sqrtm(sigma1 @ sigma2, disp=False)   # the reference's call (scipy < 1.18)
    except TypeError:
        covmean = linalg.sqrtm(sigma1 @ sigma2)                  # scipy >= 1.18 dropped `disp` and returns the matrix alone
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff @ diff + np.trace(sigma1 + sigma2 - 2 * covmean))


def fid_lowrank(real_images, generated_images):
    """The same distance without d x d matrices (float64): with a_i = (x_i - mu_i) / sqrt(n_i - 1), sigma_i = a_i^T a_i and the
    non-zero eigenvalues of sigma1 sigma2 are the squared singular values of a1 a2^T, so tr sqrtm(sigma1 sigma2) is its
    nuclear norm.  tests/test_metrics_cpu.py checks it against ``fid`` (the reference's sqrtm form) where that form is cheap;
    it is the checker at sizes where scipy's d x d sqrtm takes minutes (d = 3,072: 47 s; d = 12,288: ~1 h)."""
    x1 = np.asarray(real_images, dtype=np.float64).reshape(len(real_images), -1)
    x2 = np.asarray(generated_images, dtype=np.float64).reshape(len(generated_images), -1)
    mu1, mu2 = x1.mean(0), x2.mean(0)
    a1, a2 = (x1 - mu1) / np.sqrt(len(x1) - 1), (x2 - mu2) / np.sqrt(len(x2) - 1)
    nuc = np.linalg.svd(a1 @ a2.T, compute_uv=False).sum()
    return float(((mu1 - mu2) ** 2).sum() + (a1 ** 2).sum() + (a2 ** 2).sum() - 2 * nuc)


def _ssim_plane(x, y, data_range, win=7, k1=0.01, k2=0.03):
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    npix = win * win
    cov_norm = npix / (npix - 1)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return s[pad:-pad, pad:-pad].mean(dtype=np.float64)


def ssim(img1, img2, data_range=255):
    """img1, img2: [H, W, C] or [H, W] arrays -> mean SSIM (skimage.metrics.structural_similarity semantics, see header)."""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(img1.shape[:2]) < 7:
        raise ValueError("win_size exceeds image extent.")
    if img1.ndim == 2:
        return float(_ssim_plane(img1, img2, data_range))
    return float(np.mean([_ssim_plane(img1[..., c], img2[..., c], data_range) for c in range(img1.shape[2])]))
