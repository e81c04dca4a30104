"""One velocity evaluation between cudaProfilerStart / cudaProfilerStop, for `ncu --profile-from-start off`:

    ncu --set full --clock-control none --profile-from-start off -o /tmp/forward_full python tools/ncu_one_forward.py --mb 256
    ncu -i /tmp/forward_full.ncu-rep --page raw --csv > gpurun_out/forward_full_raw.csv
    python tools/ncu_full_summary.py gpurun_out/forward_full_raw.csv --micro-batch 256 --json profiles/r2_ncu_traffic.json > profiles/r2_ncu_forward_full.md
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rectified_flow_vision_b200 as pkg  # noqa: E402
from rectified_flow_vision_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--size", type=int, default=64)
a = ap.parse_args()
torch.manual_seed(0)
m = pkg.BaseFlowModel(image_size=a.size, device="cuda:0")
eng = E.Engine(m.velocity_net.arch(), a.size, torch.device("cuda:0"), micro_batch=a.mb, flags=0)
eng.sync_weights(m.velocity_net)
x = torch.randn(a.mb, 3, a.size, a.size, device="cuda:0")
t = torch.full((a.mb,), 0.37, device="cuda:0")
for _ in range(2):
    eng.velocity(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.velocity(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one forward profiled")
