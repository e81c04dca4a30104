"""A/B timing of the Euler loop (rfv_euler_sample, one lane) under engine flags: python tools/euler_ab.py --flags 0 262144"""
import argparse
import sys
import torch
sys.path.insert(0, ".")
from tests import util
from rectified_flow_vision_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, nargs="+", default=[0])
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--batches", type=int, nargs="+", default=[64, 256])
a = ap.parse_args()
m = util.seeded_model("default64", device="cuda:0")
for B in a.batches:
    x = torch.randn(B, 3, 64, 64, device="cuda:0")
    for fl in a.flags:
        eng = E.Engine(m.velocity_net.arch(), 64, torch.device("cuda:0"), micro_batch=B, flags=fl)
        eng.sync_weights(m.velocity_net)
        for _ in range(3):
            out, _ = eng.euler_sample(x, a.steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            out, _ = eng.euler_sample(x, a.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"batch {B} steps {a.steps} flags {fl}: {ms:.3f} ms per loop, {ms / a.steps:.3f} ms per step, "
              f"{B * a.steps / ms * 1e3:.0f} image-steps/s; checksum {float(out.double().sum()):.4f}")
