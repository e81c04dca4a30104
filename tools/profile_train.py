"""Per-op timing table of one training step (forward + backward + clip/AdamW), engine profiling mode.

    python tools/profile_train.py [--mb 128] [--size 64] [--dropout 0.1]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rectified_flow_vision_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--flags", type=int, default=0)
    a = ap.parse_args()
    torch.manual_seed(0)
    m = pkg.RectifiedFlowModel(image_size=a.size, device="cuda:0")
    os.environ["RFV_FLAGS"] = str(a.flags)
    eng = m.velocity_net.train_engine(a.size, "cuda:0", micro_batch=a.mb)
    x0 = torch.randn(a.mb, 3, a.size, a.size, device="cuda:0")
    x1 = torch.randn(a.mb, 3, a.size, a.size, device="cuda:0")
    t = torch.rand(a.mb, device="cuda:0")

    def step(i):
        eng.zero_grad()
        eng.train_accumulate(x0, x1, t, dropout_p=a.dropout, seed=i)
        return eng.optimizer_step(1e-4, i + 1)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for i in range(a.reps):
        eng.zero_grad()
        eng.train_accumulate(x0, x1, t, dropout_p=a.dropout, seed=i)
    e1.record()
    for i in range(a.reps):
        eng.optimizer_step(1e-4, i + 4)
    e2.record()
    torch.cuda.synchronize()
    fb_ms, opt_ms = e0.elapsed_time(e1) / a.reps, e1.elapsed_time(e2) / a.reps
    fl = eng.flops_per_image() * a.mb
    print(f"micro_batch={a.mb} size={a.size}: fwd+bwd {fb_ms:.3f} ms, clip+AdamW+repack {opt_ms:.3f} ms -> "
          f"{a.mb / (fb_ms + opt_ms) * 1e3:.0f} images/s; 3x-forward-FLOPs rate {3 * fl / (fb_ms + opt_ms) / 1e9:.1f} TFLOP/s; "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB (torch side only)")
    eng.set_profiling(True)
    for i in range(a.reps):
        eng.zero_grad()
        eng.train_accumulate(x0, x1, t, dropout_p=a.dropout, seed=i)
    rep = eng.profile_report()
    eng.set_profiling(False)
    rows = []
    for ln in rep.strip().splitlines():
        key, ms, n, f, _by = ln.split("\t")
        rows.append((key, float(ms) / a.reps, float(f) * a.mb))
    tot = sum(r[1] for r in rows)
    kinds = {}
    for key, ms, f in rows:
        k = key.split(" ")[0] + ("/bwd" if " bwd:" in key else "")
        v = kinds.setdefault(k, [0.0, 0.0])
        v[0] += ms
        v[1] += f
    for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:<24}{v[0]:>9.3f} ms {v[0] / tot:>7.3f}" + (f"{v[1] / v[0] / 1e9:>9.0f} TF/s" if v[1] > 0 else ""))
    print(f"summed {tot:.3f} ms")
    print(f"{'op':<60}{'ms':>9}{'share':>8}{'TFLOP/s':>10}")
    for key, ms, f in sorted(rows, key=lambda r: -r[1])[:a.top]:
        print(f"{key:<60}{ms:>9.4f}{ms / tot:>8.3f}{f / ms / 1e9 if ms > 0 else 0:>10.1f}")


if __name__ == "__main__":
    main()
