"""GPU-side quality metrics: the MetricsCalculator surface of the reference (utils/metrics.py:17-173) on device tensors.

The reference flattens the images to numpy, takes ``np.mean`` / ``np.cov`` (a d x d float64 matrix, d = 12,288 at 64x64),
``scipy.linalg.sqrtm`` of a d x d product, and ``skimage.metrics.structural_similarity`` one image pair at a time.  Here
the tensors stay in HBM: mean / covariance / SSIM are CUDA kernels behind the C ABI (``rfv_metrics_*``, csrc/metrics.cuh),
and the Frechet distance is evaluated from an n1 x n2 cross Gram matrix whose singular values sum to
tr sqrtm(sigma1 sigma2) -- the d x d matrices are only formed when ``compute_fid_statistics`` is asked for them.
No CPU fallback: the calls raise if the CUDA library is missing.
"""
from __future__ import annotations

import time
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import engine as E


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _sync():
    if torch.cuda.is_available():   # (the timing loop itself is host logic; utils/metrics.py:150,158 guards the same way)
        torch.cuda.synchronize()


def fid_statistics(images: torch.Tensor, covariance: bool = True):
    """images [N, ...] on a CUDA device -> (mu [d] fp64, sigma [d, d] fp64 or None), both device tensors."""
    if images.dim() < 2 or images.shape[0] < 1:
        raise ValueError("images: expected [N, ...] with N >= 1")
    dev = images.device
    x = _dev_f32(images, dev).reshape(images.shape[0], -1)
    n, d = x.shape
    lib = E.load_library()
    mu = torch.empty(d, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        E._check(lib.rfv_metrics_mean(x.data_ptr(), n, d, mu.data_ptr(), _stream(dev)))
        sigma = None
        if covariance:
            if n < 2:
                raise ValueError("covariance needs at least two images")
            sigma = torch.empty((d, d), dtype=torch.float64, device=dev)
            E._check(lib.rfv_metrics_covariance(x.data_ptr(), mu.data_ptr(), n, d, sigma.data_ptr(), _stream(dev)))
    return mu, sigma


def fid(real_images: torch.Tensor, generated_images: torch.Tensor) -> float:
    """Frechet distance between the pixel-space Gaussians of two image sets (utils/metrics.py:89-116), device tensors in."""
    if real_images.shape[1:] != generated_images.shape[1:]:
        raise ValueError("image sets must share the per-image shape")
    dev = real_images.device
    x1 = _dev_f32(real_images, dev).reshape(real_images.shape[0], -1)
    x2 = _dev_f32(generated_images, dev).reshape(generated_images.shape[0], -1)
    (n1, d), n2 = x1.shape, x2.shape[0]
    if n1 < 2 or n2 < 2:
        raise ValueError("FID needs at least two images per set")
    lib = E.load_library()
    mu1 = torch.empty(d, dtype=torch.float64, device=dev)
    mu2 = torch.empty(d, dtype=torch.float64, device=dev)
    gram = torch.empty((n1, n2), dtype=torch.float64, device=dev)
    terms = torch.empty(3, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        s = _stream(dev)
        E._check(lib.rfv_metrics_mean(x1.data_ptr(), n1, d, mu1.data_ptr(), s))
        E._check(lib.rfv_metrics_mean(x2.data_ptr(), n2, d, mu2.data_ptr(), s))
        E._check(lib.rfv_metrics_fid_terms(x1.data_ptr(), mu1.data_ptr(), n1, x2.data_ptr(), mu2.data_ptr(), n2, d,
                                           gram.data_ptr(), terms.data_ptr(), s))
    t = terms.cpu().numpy()
    # tr sqrtm(sigma1 sigma2) = nuclear norm of the cross Gram matrix (csrc/metrics.cuh); an n1 x n2 SVD on the host
    nuc = float(np.linalg.svd(gram.cpu().numpy(), compute_uv=False).sum())
    return float(t[0] + t[1] + t[2] - 2.0 * nuc)


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 255.0) -> torch.Tensor:
    """Mean SSIM per image pair: a, b [B, C, H, W] device tensors -> [B] fp64 device tensor (skimage defaults)."""
    if a.shape != b.shape:
        raise ValueError("Images must have the same size")
    if a.dim() != 4:
        raise ValueError("expected [B, C, H, W]")
    bsz, c, h, w = a.shape
    if h < 7 or w < 7:
        raise ValueError("win_size exceeds image extent")
    dev = a.device
    a32, b32 = _dev_f32(a, dev), _dev_f32(b, dev)
    out = torch.empty(bsz, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        E._check(E.load_library().rfv_metrics_ssim(a32.data_ptr(), b32.data_ptr(), bsz, c, h, w, float(data_range),
                                                   out.data_ptr(), _stream(dev)))
    return out


class MetricsCalculator:
    """Call-compatible with utils/metrics.py:17-173; the arithmetic runs on ``device`` (a CUDA device)."""

    def __init__(self, device: str = "cuda"):
        self.device = device
        self._lpips_model = None
        self._inception_model = None

    @property
    def lpips_model(self):
        """Lazy LPIPS network (third-party package + AlexNet weights; absent in this image -> None, as in the reference)."""
        if self._lpips_model is None:
            try:
                import lpips
                self._lpips_model = lpips.LPIPS(net="alex").to(self.device)
                self._lpips_model.eval()
            except ImportError:
                print("LPIPS not available. Install with: pip install lpips")
                return None
        return self._lpips_model

    def compute_ssim(self, img1: np.ndarray, img2: np.ndarray) -> float:
        """img1, img2: [H, W, C] (or [H, W]) arrays in [0, 255] -> SSIM (utils/metrics.py:39-53)."""
        if img1.shape != img2.shape:
            raise ValueError("Images must have the same size")
        def chw(im):
            t = torch.as_tensor(np.ascontiguousarray(im), dtype=torch.float32)
            t = t.permute(2, 0, 1) if t.dim() == 3 else t.unsqueeze(0)
            return t.unsqueeze(0).to(self.device)
        return float(ssim(chw(img1), chw(img2), 255.0).item())

    def compute_lpips(self, img1: torch.Tensor, img2: torch.Tensor) -> float:
        if self.lpips_model is None:
            return float("nan")
        with torch.no_grad():
            return self.lpips_model(img1.to(self.device), img2.to(self.device)).mean().item()

    def compute_fid_statistics(self, images: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
        """images [N, C, H, W] -> (mu [d] float32, sigma [d, d] float64) as numpy arrays (utils/metrics.py:73-87)."""
        mu, sigma = fid_statistics(images.to(self.device))
        return mu.to(torch.float32).cpu().numpy(), sigma.cpu().numpy()

    def compute_fid(self, real_images: torch.Tensor, generated_images: torch.Tensor) -> float:
        return fid(real_images.to(self.device), generated_images.to(self.device))

    def compute_generation_speed(self, model, num_samples: int, num_steps: int, batch_size: int = 1, num_runs: int = 5,
                                 image_size: int = 64) -> Dict[str, float]:
        """Wall-clock generation speed of ``model.sample`` (utils/metrics.py:118-173): same loop, same result keys."""
        model.eval()
        times = []
        with torch.no_grad():
            for run in range(num_runs):
                if run == 0:
                    model.sample(torch.randn(1, 3, image_size, image_size).to(self.device), num_steps=num_steps)
                _sync()
                t0 = time.time()
                for i in range(0, num_samples, batch_size):
                    cur = min(batch_size, num_samples - i)
                    model.sample(torch.randn(cur, 3, image_size, image_size).to(self.device), num_steps=num_steps)
                _sync()
                times.append(time.time() - t0)
        total = float(np.mean(times))
        return {"total_time": total, "time_per_image": total / num_samples, "images_per_second": num_samples / total,
                "time_std": float(np.std(times)), "num_steps": num_steps, "num_samples": num_samples}


def benchmark_models(base_model, rectified_model, steps_list: List[int], num_samples: int = 50, image_size: int = 64,
                     device: str = "cuda") -> Dict:
    """Generation speed of a base and a rectified model over a list of step counts (utils/metrics.py:175-222): same
    arguments, same result dictionary ({'base_model': [...], 'rectified_model': [...]}, one speed record per step count
    with 'model' added), same console report."""
    calc = MetricsCalculator(device)
    results = {"base_model": [], "rectified_model": []}
    print("\n" + "=" * 60)
    print("BENCHMARK: Modelo Base vs Modelo Rectificado")
    print("=" * 60)
    for num_steps in steps_list:
        rec = {}
        for key, tag, model in (("base_model", "base", base_model), ("rectified_model", "rectified", rectified_model)):
            rec[tag] = calc.compute_generation_speed(model, num_samples, num_steps, image_size=image_size)
            rec[tag]["model"] = tag
            results[key].append(rec[tag])
        print(f"\nPasos: {num_steps}")
        print(f"  Base:       {rec['base']['time_per_image'] * 1000:.2f} ms/img")
        print(f"  Rectified:  {rec['rectified']['time_per_image'] * 1000:.2f} ms/img")
    return results
