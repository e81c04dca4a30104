"""Golden values for the quality metrics from the reference's OWN MetricsCalculator (run in the build container only).

    python oracle/make_golden_metrics.py      # writes tests/golden/metrics.npz

`utils/metrics.py` of the reference is loaded as a file (``utils/__init__.py`` imports plotting modules that are not
installed) with two shims, both recorded here because they decide what "the reference's output" means in this container:
  * ``skimage`` is absent: a stub module satisfies the import at utils/metrics.py:10.  compute_fid_statistics / compute_fid
    (utils/metrics.py:73-116) never touch it, so their goldens ARE the unmodified reference's results; the SSIM goldens come
    from oracle/metrics_oracle.ssim (the restatement of skimage 0.21's algorithm) and are marked as such in the file;
  * scipy >= 1.18 dropped the ``disp`` argument the reference passes to ``scipy.linalg.sqrtm`` (utils/metrics.py:107; the
    reference pins scipy==1.11.3): the call is routed through a wrapper that returns ``(sqrtm(a), 0.0)`` as old scipy did.
Inputs are seeded numpy draws stored next to the outputs.  TEST INFRASTRUCTURE: nothing in the product path imports this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as M  # noqa: E402

REF = "/root/reference"


def load_reference_metrics():
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.metrics")
    skm.structural_similarity = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("skimage is not installed"))
    sk.metrics = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.metrics", skm)
    import scipy.linalg as sl
    real = sl.sqrtm

    def sqrtm_compat(a, disp=True, **kw):
        try:
            return real(a, disp=disp, **kw)
        except TypeError:
            r = real(a, **kw)
            return r if disp else (r, 0.0)

    sl.sqrtm = sqrtm_compat
    spec = importlib.util.spec_from_file_location("ref_utils_metrics", os.path.join(REF, "utils", "metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_metrics()
    calc = ref.MetricsCalculator(device="cpu")
    rng = np.random.default_rng(20261019)
    out = {}
    # FID statistics and FID through the reference's own methods (torch tensors in, as its callers pass them)
    x_stats = (rng.standard_normal((10, 3, 8, 8)) * 1.5 + 0.3).astype(np.float32)
    mu, sigma = calc.compute_fid_statistics(torch.from_numpy(x_stats))
    out.update(x_stats=x_stats, mu=mu, sigma=sigma)
    cases = [(300, 280, (3, 6, 6)), (10, 12, (3, 8, 8)), (40, 7, (1, 9, 9))]
    fids = []
    for i, (n1, n2, shape) in enumerate(cases):
        a = rng.standard_normal((n1,) + shape).astype(np.float32)
        b = (rng.standard_normal((n2,) + shape) * 1.7 + 0.5).astype(np.float32)
        out[f"fid_a{i}"], out[f"fid_b{i}"] = a, b
        fids.append(calc.compute_fid(torch.from_numpy(a), torch.from_numpy(b)))
        assert abs(fids[-1] - M.fid(a, b)) <= 1e-9 * abs(fids[-1])   # the oracle's fid IS these calls
    out["fid"] = np.array(fids)
    out["fid_identical"] = np.array([calc.compute_fid(torch.from_numpy(out["fid_a1"]), torch.from_numpy(out["fid_a1"]))])
    # SSIM: restatement (skimage absent) -- flagged
    ss = []
    for i, shape in enumerate([(64, 64, 3), (40, 56, 3), (7, 7, 1), (33, 50)]):
        x = rng.integers(0, 256, shape).astype(np.uint8)
        y = np.clip(x.astype(np.float64) + rng.normal(0, 25, shape), 0, 255).astype(np.uint8)
        out[f"ssim_x{i}"], out[f"ssim_y{i}"] = x, y
        ss.append(M.ssim(x, y))
    out["ssim"] = np.array(ss)
    out["ssim_source"] = np.array("oracle/metrics_oracle.ssim (restatement of skimage 0.21; skimage not installed)")
    out["fid_source"] = np.array("unmodified /root/reference utils/metrics.py MetricsCalculator (skimage import stubbed, sqrtm disp shim)")
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: fid {fids}, identical {float(out['fid_identical'][0]):.3e}, ssim {ss}")


if __name__ == "__main__":
    main()
