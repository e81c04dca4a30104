"""Golden benchmark_report.txt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden_report.py     # writes tests/golden/benchmark_report.{json,txt}

utils/visualization.py:210-253 (create_summary_report) of the reference writes the text report the package's
``benchmark.create_summary_report`` has to reproduce byte for byte.  The module imports matplotlib / seaborn / PIL at the
top, none of which is installed here, and the function ends by drawing a plot: those three packages are replaced by inert
stand-ins so that the reference's own text-writing code runs unmodified.

TEST INFRASTRUCTURE: nothing in the product path imports this.
"""
import json
import os
import sys
import tempfile
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

RESULTS = {
    "base_model": [{"num_steps": n, "total_time": t, "time_per_image": t / 64, "images_per_second": 64 / t, "time_std": 0.01,
                    "num_samples": 64} for n, t in ((1, 0.0517), (2, 0.1093), (4, 0.2011), (8, 0.4102), (100, 5.0371))],
    "rectified_model": [{"num_steps": n, "total_time": t, "time_per_image": t / 64, "images_per_second": 64 / t, "time_std": 0.02,
                         "num_samples": 64} for n, t in ((1, 0.0498), (2, 0.0777), (4, 0.2222), (8, 0.3999), (100, 4.9))],
}


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "PIL", "PIL.Image"):
        sys.modules[name] = mock.MagicMock(name=name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.pyplot"].subplots.return_value = (mock.MagicMock(), (mock.MagicMock(), mock.MagicMock()))
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_visualization", "/root/reference/utils/visualization.py")
    vis = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vis)
    with tempfile.TemporaryDirectory() as d:
        vis.create_summary_report(RESULTS, d)
        text = open(os.path.join(d, "benchmark_report.txt")).read()
    with open(os.path.join(GOLD, "benchmark_report.txt"), "w") as f:
        f.write(text)
    with open(os.path.join(GOLD, "benchmark_report.json"), "w") as f:
        json.dump(RESULTS, f, indent=1)
    print(text)


if __name__ == "__main__":
    main()
