"""Board power and SM clock drawn by each kernel class of the forward pass when it alone runs back to back for a few seconds
(RFV_ONLY_KIND, a diagnosis switch of the engine: results are meaningless, launches and data sizes are the real ones).
Explains what the power cap (sw_power_cap) is spent on.   python tools/power_by_kind.py [--mb 256] [--seconds 3]"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

KINDS = ["", "conv_halo", "conv_umma", "gn_apply", "input_conv", "output_conv", "attention"]


def child(kind, mb, seconds):
    import torch
    sys.path.insert(0, ".")
    from tests import util
    from rectified_flow_vision_b200 import engine as E
    m = util.seeded_model("default64", device="cuda:0")
    eng = E.Engine(m.velocity_net.arch(), 64, torch.device("cuda:0"), micro_batch=mb)
    eng.sync_weights(m.velocity_net)
    x = torch.randn(mb, 3, 64, 64, device="cuda:0")
    t = torch.rand(mb, device="cuda:0")
    for _ in range(5):
        eng.velocity(x, t)
    torch.cuda.synchronize()
    fd, path = tempfile.mkstemp(suffix=".csv")
    os.close(fd)
    smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            eng.velocity(x, t)
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    smi.terminate()
    smi.wait(timeout=5)
    rows = [[c.strip() for c in l.split(",")] for l in open(path) if l.count(",") >= 2]
    rows = rows[len(rows) // 3:]   # steady state
    pw = sorted(float(r[0]) for r in rows)
    ck = sorted(float(r[1]) for r in rows)
    cap = sum(1 for r in rows if "Active" in r[2] and "Not" not in r[2])
    print(f"{kind or 'whole forward':<14} {ms:8.3f} ms/call  power {pw[len(pw) // 2]:7.1f} W  sm {ck[len(ck) // 2]:6.0f} MHz  "
          f"power-capped samples {cap}/{len(rows)}  energy {ms * 1e-3 * pw[len(pw) // 2]:.3f} J/call", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--kind", default=None)
    a = ap.parse_args()
    if a.kind is not None:
        child(a.kind, a.mb, a.seconds)
    else:
        for k in KINDS:
            env = dict(os.environ)
            env["RFV_ONLY_KIND"] = k
            subprocess.run([sys.executable, __file__, "--mb", str(a.mb), "--seconds", str(a.seconds), "--kind", k], env=env, check=False)
