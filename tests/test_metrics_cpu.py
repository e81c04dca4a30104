"""CPU checks of the quality-metric oracle (oracle/metrics_oracle.py) against the reference's own tests
(tests/test_utils.py:30-73 of the reference) and closed forms; no GPU needed."""
import numpy as np
import pytest

import os

from oracle import metrics_oracle as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


def test_oracle_reproduces_the_reference_goldens():
    """tests/golden/metrics.npz: FID statistics / FID computed by the reference's OWN MetricsCalculator
    (oracle/make_golden_metrics.py loads utils/metrics.py unmodified, skimage import stubbed, sqrtm `disp` shim)."""
    g = np.load(GOLD)
    mu, sigma = M.fid_statistics(g["x_stats"])
    assert np.array_equal(mu.astype(np.float32), g["mu"]) and np.allclose(sigma, g["sigma"], rtol=1e-12, atol=1e-14)
    for i, want in enumerate(g["fid"]):
        got = M.fid(g[f"fid_a{i}"], g[f"fid_b{i}"])
        assert abs(got - want) <= 1e-9 * abs(want), (i, got, want)
        assert abs(M.fid_lowrank(g[f"fid_a{i}"], g[f"fid_b{i}"]) - want) <= 1e-6 * abs(want)
    for i, want in enumerate(g["ssim"]):   # (these four come from the restatement itself: skimage is absent)
        assert abs(M.ssim(g[f"ssim_x{i}"], g[f"ssim_y{i}"]) - want) <= 1e-14


def test_ssim_reference_test_cases():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (64, 64, 3), dtype=np.uint8)
    assert M.ssim(img, img) > 0.99                      # tests/test_utils.py:30-34
    assert abs(M.ssim(img, img) - 1.0) < 1e-12
    z, w = np.zeros((64, 64, 3), np.uint8), np.ones((64, 64, 3), np.uint8) * 255
    assert M.ssim(z, w) < 0.5                           # tests/test_utils.py:36-41
    c1 = (0.01 * 255) ** 2
    assert abs(M.ssim(z, w) - c1 / (255.0 ** 2 + c1)) < 1e-15   # constant images: only the luminance term is left
    with pytest.raises(ValueError):                     # tests/test_utils.py:43-49
        M.ssim(img, img[:32, :32])


def test_ssim_constant_images_closed_form_and_symmetry():
    a, b = np.full((16, 20), 40.0), np.full((16, 20), 200.0)
    c1 = (0.01 * 255) ** 2
    assert abs(M.ssim(a, b) - (2 * 40 * 200 + c1) / (40 ** 2 + 200 ** 2 + c1)) < 1e-14
    rng = np.random.default_rng(1)
    x, y = rng.integers(0, 256, (24, 31, 3)).astype(np.uint8), rng.integers(0, 256, (24, 31, 3)).astype(np.uint8)
    assert abs(M.ssim(x, y) - M.ssim(y, x)) < 1e-15
    assert -1.0 <= M.ssim(x, y) <= 1.0


def test_fid_reference_test_cases():
    rng = np.random.default_rng(2)
    images = rng.standard_normal((10, 3, 8, 8)).astype(np.float32)
    mu, sigma = M.fid_statistics(images)                # tests/test_utils.py:51-58 (shape contract)
    assert mu.shape == (192,) and sigma.shape == (192, 192) and sigma.dtype == np.float64
    assert M.fid(images, images) < 1.0                  # tests/test_utils.py:60-65
    other = (rng.standard_normal((10, 3, 8, 8)) * 2 + 1).astype(np.float32)
    assert M.fid(images, other) > 0                     # tests/test_utils.py:67-73


@pytest.mark.parametrize("n1,n2,shape", [(300, 280, (3, 6, 6)), (10, 12, (3, 8, 8)), (40, 7, (1, 9, 9))])
def test_nuclear_norm_form_equals_the_sqrtm_form(n1, n2, shape):
    """The identity the CUDA path rests on: tr sqrtm(sigma1 sigma2) = sum of the singular values of the cross Gram matrix."""
    rng = np.random.default_rng(n1)
    x1 = rng.standard_normal((n1,) + shape).astype(np.float32)
    x2 = (rng.standard_normal((n2,) + shape) * 1.7 + 0.5).astype(np.float32)
    ref, low = M.fid(x1, x2), M.fid_lowrank(x1, x2)
    assert abs(ref - low) <= 1e-6 * abs(ref), (ref, low)
    assert abs(M.fid_lowrank(x1, x1)) < 1e-9


class _StubModel:
    """Host-logic stand-in for a flow model: counts what the timing loops ask of it."""

    def __init__(self):
        self.calls = []

    def eval(self):
        return self

    def sample(self, noise, num_steps=100):
        self.calls.append((tuple(noise.shape), num_steps))
        return noise


def test_generation_speed_and_benchmark_models_host_logic(capsys):
    """compute_generation_speed / benchmark_models (utils/metrics.py:118-222): batching, warm-up call, result keys."""
    from rectified_flow_vision_b200.metrics import MetricsCalculator, benchmark_models
    m = _StubModel()
    r = MetricsCalculator(device="cpu").compute_generation_speed(m, num_samples=5, num_steps=3, batch_size=2, num_runs=2, image_size=8)
    assert set(r) == {"total_time", "time_per_image", "images_per_second", "time_std", "num_steps", "num_samples"}
    assert r["num_steps"] == 3 and r["num_samples"] == 5 and r["images_per_second"] > 0
    # one warm-up sample of one image, then per run batches of 2, 2, 1
    assert m.calls == [((1, 3, 8, 8), 3)] + [((2, 3, 8, 8), 3), ((2, 3, 8, 8), 3), ((1, 3, 8, 8), 3)] * 2
    b, q = _StubModel(), _StubModel()
    res = benchmark_models(b, q, steps_list=[1, 4], num_samples=2, image_size=8, device="cpu")
    assert [x["num_steps"] for x in res["base_model"]] == [1, 4] and [x["model"] for x in res["rectified_model"]] == ["rectified"] * 2
    out = capsys.readouterr().out
    assert "BENCHMARK: Modelo Base vs Modelo Rectificado" in out and "Pasos: 4" in out
