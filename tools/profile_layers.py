"""Per-op timing table of one velocity evaluation (CUDA events around every launch; engine profiling mode).

    python tools/profile_layers.py [--mb 256] [--size 64] [--flags 0]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rectified_flow_vision_b200 as pkg  # noqa: E402
from rectified_flow_vision_b200 import engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="images per call (default: the micro-batch)")
    a = ap.parse_args()
    torch.manual_seed(0)
    m = pkg.BaseFlowModel(image_size=a.size, device="cuda:0")
    eng = E.Engine(m.velocity_net.arch(), a.size, torch.device("cuda:0"), micro_batch=a.mb, flags=a.flags)
    eng.sync_weights(m.velocity_net)
    nb = a.batch or a.mb
    x = torch.randn(nb, 3, a.size, a.size, device="cuda:0")
    for _ in range(3):
        eng.euler_sample(x, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.euler_sample(x, 1)
    e1.record()
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / a.reps
    eng.set_profiling(True)
    for _ in range(a.reps):
        eng.euler_sample(x, 1)
    rep = eng.profile_report()
    eng.set_profiling(False)
    rows = []
    for ln in rep.strip().splitlines():
        key, ms, n, fl, _by = ln.split("\t")
        rows.append((key, float(ms) / a.reps, float(fl) * nb))
    tot = sum(r[1] for r in rows)
    kinds = {}
    for key, ms, fl in rows:
        k = kinds.setdefault(key.split(" ")[0], [0.0, 0.0])
        k[0] += ms
        k[1] += fl
    print("  ".join(f"{k}={v[0]:.3f}ms" + (f"({v[1] / v[0] / 1e9:.0f}TF/s)" if v[1] > 0 else "") for k, v in sorted(kinds.items())))
    print(f"micro_batch={a.mb} batch={nb} size={a.size} flags={a.flags}: forward {fwd_ms:.3f} ms unprofiled, {tot:.3f} ms summed; "
          f"{nb / fwd_ms * 1e3:.0f} img-steps/s; {eng.flops_per_image() * nb / fwd_ms / 1e9:.1f} TFLOP/s")
    print(f"{'op':<52}{'ms':>9}{'share':>8}{'TFLOP/s':>10}")
    order = {op: i for i, op in enumerate(k for k, _, _ in rows)}
    for key, ms, fl in sorted(rows, key=lambda r: -r[1]):
        print(f"{key:<52}{ms:>9.4f}{ms / tot:>8.3f}{fl / ms / 1e9 if ms > 0 else 0:>10.1f}")


if __name__ == "__main__":
    main()
