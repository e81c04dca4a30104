"""GPU-side quality metrics (csrc/metrics.cuh behind rfv_metrics_*) against the CPU oracle (oracle/metrics_oracle.py), the
reference's own test cases (tests/test_utils.py:30-73) and size-independent properties at full size."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


def test_against_goldens_from_the_reference_calculator():
    """tests/golden/metrics.npz (oracle/make_golden_metrics.py): statistics and FID from the reference's own
    MetricsCalculator.compute_fid_statistics / compute_fid on seeded inputs; SSIM from the skimage restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz"))
    c = _calc()
    mu, sigma = c.compute_fid_statistics(torch.from_numpy(g["x_stats"]))
    assert np.abs(mu - g["mu"]).max() <= 1e-6
    assert np.abs(sigma - g["sigma"]).max() <= 2e-6 * np.abs(g["sigma"]).max()
    for i, want in enumerate(g["fid"]):
        got = c.compute_fid(torch.from_numpy(g[f"fid_a{i}"]), torch.from_numpy(g[f"fid_b{i}"]))
        assert abs(got - want) <= 2e-6 * abs(want), (i, got, want)
    same = c.compute_fid(torch.from_numpy(g["fid_a1"]), torch.from_numpy(g["fid_a1"]))
    assert abs(same) < 1e-3 and abs(float(g["fid_identical"][0])) < 1e-3      # both ~0 (the reference's sqrtm: -3e-5)
    for i, want in enumerate(g["ssim"]):
        assert abs(c.compute_ssim(g[f"ssim_x{i}"], g[f"ssim_y{i}"]) - want) <= 1e-10


def _calc():
    from rectified_flow_vision_b200.metrics import MetricsCalculator
    return MetricsCalculator(device="cuda:0")


@pytest.mark.parametrize("n,shape", [(10, (3, 32, 32)), (2, (3, 8, 8)), (65, (1, 7, 19)), (257, (3, 5, 5))])
def test_fid_statistics_vs_oracle(n, shape):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((n,) + shape) * 1.5 + 0.3).astype(np.float32)
    mu, sigma = _calc().compute_fid_statistics(torch.from_numpy(x))
    mu_ref, sigma_ref = M.fid_statistics(x)
    d = int(np.prod(shape))
    assert mu.shape == (d,) and sigma.shape == (d, d) and sigma.dtype == np.float64     # tests/test_utils.py:51-58
    assert np.abs(mu - mu_ref).max() <= 1e-6
    assert np.abs(sigma - sigma_ref).max() <= 2e-6 * np.abs(sigma_ref).max()            # fp32 products, fp64 accumulation
    assert np.abs(sigma - sigma.T).max() <= 1e-12 * np.abs(sigma).max()


@pytest.mark.parametrize("n1,n2,shape", [(300, 280, (3, 6, 6)), (10, 12, (3, 8, 8)), (40, 7, (1, 9, 9))])
def test_fid_vs_the_reference_sqrtm_form(n1, n2, shape):
    rng = np.random.default_rng(n1 + n2)
    x1 = rng.standard_normal((n1,) + shape).astype(np.float32)
    x2 = (rng.standard_normal((n2,) + shape) * 1.7 + 0.5).astype(np.float32)
    got = _calc().compute_fid(torch.from_numpy(x1), torch.from_numpy(x2))
    ref = M.fid(x1, x2)
    assert abs(got - ref) <= 2e-6 * abs(ref), (got, ref)


def test_fid_reference_test_cases_and_larger_sets():
    c = _calc()
    g = torch.Generator().manual_seed(5)
    images = torch.randn(10, 3, 32, 32, generator=g)
    assert c.compute_fid(images, images) < 1.0                       # tests/test_utils.py:60-65
    other = torch.randn(10, 3, 32, 32, generator=g) * 2 + 1
    f = c.compute_fid(images, other)
    assert f > 0                                                     # tests/test_utils.py:67-73
    ref = M.fid_lowrank(images.numpy(), other.numpy())               # (scipy's 3072 x 3072 sqrtm takes 47 s)
    assert abs(f - ref) <= 2e-6 * abs(ref), (f, ref)
    # full size: 2048 images of 3 x 64 x 64 (d = 12,288) -- symmetry, identity, and the float64 checker
    a = torch.randn(2048, 3, 64, 64, generator=g)
    b = torch.randn(1500, 3, 64, 64, generator=g) * 1.1 + 0.05
    fab, fba = c.compute_fid(a, b), c.compute_fid(b, a)
    assert abs(fab - fba) <= 1e-7 * abs(fab)
    assert abs(c.compute_fid(a, a)) <= 1e-6 * 12288
    ref = M.fid_lowrank(a.numpy(), b.numpy())
    assert abs(fab - ref) <= 5e-6 * abs(ref), (fab, ref)


def test_ssim_vs_oracle_and_reference_test_cases():
    c = _calc()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (64, 64, 3), dtype=np.uint8)
    assert c.compute_ssim(img, img) > 0.99                           # tests/test_utils.py:30-34
    assert abs(c.compute_ssim(img, img) - 1.0) < 1e-12
    z, w = np.zeros((64, 64, 3), np.uint8), np.ones((64, 64, 3), np.uint8) * 255
    assert c.compute_ssim(z, w) < 0.5                                # tests/test_utils.py:36-41
    assert abs(c.compute_ssim(z, w) - M.ssim(z, w)) < 1e-15
    with pytest.raises(ValueError):                                  # tests/test_utils.py:43-49
        c.compute_ssim(img, img[:32, :32])
    with pytest.raises(ValueError):
        c.compute_ssim(img[:6, :6], img[:6, :6])                     # window larger than the image
    for shape in [(64, 64, 3), (128, 96, 3), (7, 7, 1), (33, 250), (9, 64, 4)]:
        x = rng.integers(0, 256, shape).astype(np.uint8)
        noise = rng.normal(0, 25, shape)
        y = np.clip(x.astype(np.float64) + noise, 0, 255).astype(np.uint8)
        got, ref = c.compute_ssim(x, y), M.ssim(x, y)
        assert abs(got - ref) <= 1e-10, (shape, got, ref)


def test_ssim_batch_matches_per_pair_and_is_one_on_identical_batches():
    from rectified_flow_vision_b200 import metrics
    g = torch.Generator().manual_seed(9)
    a = (torch.rand(512, 3, 64, 64, generator=g) * 255).round()
    b = (a + torch.randn(512, 3, 64, 64, generator=g) * 20).clamp(0, 255).round()
    s = metrics.ssim(a.cuda(), b.cuda()).cpu().numpy()
    assert s.shape == (512,) and np.isfinite(s).all()
    for i in (0, 17, 511):
        ref = M.ssim(a[i].permute(1, 2, 0).numpy(), b[i].permute(1, 2, 0).numpy())
        assert abs(s[i] - ref) <= 1e-10
    one = metrics.ssim(a.cuda(), a.cuda()).cpu().numpy()
    assert np.abs(one - 1.0).max() < 1e-12


def test_c_abi_error_paths():
    from rectified_flow_vision_b200 import engine as E
    lib = E.load_library()
    x = torch.zeros(4, 8, device="cuda:0")
    mu = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    sig = torch.zeros(8, 8, dtype=torch.float64, device="cuda:0")
    assert lib.rfv_metrics_mean(None, 4, 8, mu.data_ptr(), None) == -1
    assert lib.rfv_metrics_covariance(x.data_ptr(), mu.data_ptr(), 1, 8, sig.data_ptr(), None) == -1   # one sample
    assert b"two samples" in lib.rfv_last_error()
    assert lib.rfv_metrics_ssim(x.data_ptr(), x.data_ptr(), 1, 1, 4, 8, 255.0, mu.data_ptr(), None) == -1
    assert b"window" in lib.rfv_last_error()
