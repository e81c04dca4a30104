"""Device timing of the GPU-side quality metrics (CUDA events): FID terms of 2048 x 2048 images of 3x64x64, covariance, SSIM of 512 pairs."""
import sys
import time
import torch
sys.path.insert(0, ".")
from rectified_flow_vision_b200 import metrics  # noqa: E402

g = torch.Generator().manual_seed(0)
a = torch.randn(2048, 3, 64, 64, generator=g).cuda()
b = (torch.randn(2048, 3, 64, 64, generator=g) * 1.1).cuda()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t0 = time.time()
f = metrics.fid(a, b)
print(f"fid(2048, 2048 x 12288) = {f:.4f}; end to end incl. the host SVD: {(time.time() - t0) * 1e3:.0f} ms (first call)")
t0 = time.time()
f = metrics.fid(a, b)
print(f"   second call: {(time.time() - t0) * 1e3:.0f} ms")
print(f"mean + covariance (2048 x 12288 -> 12288 x 12288 fp64): {timed(lambda: metrics.fid_statistics(a)):.2f} ms "
      f"({2.0 * 12288 * 12288 * 2048 / 1e9:.0f} GFLOP fp32 FMA)")
x = (torch.rand(512, 3, 64, 64, generator=g) * 255).round().cuda()
y = (x + torch.randn(512, 3, 64, 64, device="cuda") * 20).clamp(0, 255).round()
ms = timed(lambda: metrics.ssim(x, y), 10)
print(f"ssim of 512 pairs of 3x64x64: {ms:.3f} ms ({512 / ms * 1e3:.0f} pairs/s)")
