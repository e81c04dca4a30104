"""Same-box A/B of the production pair-generation path under environment switches (one process per variant, alternated):
    python tools/pairs_ab.py RFV_WA_TMEM=1 RFV_WA_TMEM=2 ...      ("-" = no switch)
Each variant integrates 2048 noises for 100 Euler steps (micro-batch 512, two lanes, loop graphs) `--reps` times."""
import argparse
import os
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from tests import util
m = util.seeded_model("default64", device="cuda:0")
eng = m._engine(64)
x = torch.randn(2048, 3, 64, 64, device="cuda:0")
eng.euler_sample(x, 20)
torch.cuda.synchronize()
reps = int(sys.argv[1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.euler_sample(x, 100)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{2048 / ms * 1e3:.1f} pairs/s ({ms:.1f} ms per 2048 pairs)")
'''

ap = argparse.ArgumentParser()
ap.add_argument("variants", nargs="+")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--rounds", type=int, default=2)
a = ap.parse_args()
for r in range(a.rounds):
    for v in a.variants:
        env = dict(os.environ)
        if v != "-":
            for kv in v.split(","):
                k, val = kv.split("=")
                env[k] = val
        out = subprocess.run([sys.executable, "-c", CHILD, str(a.reps)], env=env, capture_output=True, text=True)
        print(f"round {r} {v:28s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
