"""Layer-by-layer parity diagnostic on a GPU box (uses the oracle as checker; not part of the product path).

    python tools/gpu_diag.py [--no-umma] [--case small32|default64]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from tests import util  # noqa: E402
from rectified_flow_vision_b200 import engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-umma", action="store_true")
    ap.add_argument("--case", default="small32")
    ap.add_argument("--micro-batch", type=int, default=4)
    ap.add_argument("--flags", type=int, default=0, help="extra RFV_FLAG_* bits")
    a = ap.parse_args()
    flags = 4 | (1 if a.no_umma else 0) | a.flags
    m = util.seeded_model(a.case)
    g = util.golden(a.case)
    spec = util.spec_for(a.case)
    P = util.numpy_params(m)
    taps = {}
    t0 = time.time()
    v_or = O.unet_forward(P, g["x"], g["t"], spec, policy=O.BF16_POLICY, taps=taps)
    print(f"oracle(bf16 policy) {time.time()-t0:.1f}s; vs golden fp32 rel_l2={util.rel_l2(v_or, g['v']):.4e}")
    dev = torch.device("cuda:0")
    m.device = "cuda:0"
    m.to(dev)
    size = g["x"].shape[-1]
    eng = E.Engine(m.velocity_net.arch(), size, dev, micro_batch=a.micro_batch, flags=flags)
    eng.sync_weights(m.velocity_net)
    torch.cuda.synchronize()
    # weight pack round trip
    w = m.velocity_net.enc_blocks._modules["0"].conv1.weight.detach()
    back = eng.get_tensor("velocity_net.enc_blocks.0.conv1.weight", w.numel()).view_as(w)
    print("pack roundtrip max|d| vs bf16(w):", float((back - w.bfloat16().float()).abs().max()))
    x = torch.from_numpy(g["x"]).to(dev)
    t = torch.from_numpy(g["t"]).to(dev)
    v = eng.velocity(x, t)
    torch.cuda.synchronize()
    v = v.cpu().numpy()
    B = x.shape[0]
    print(f"{'layer':<18}{'rel_l2':>12}{'max_rel':>12}{'rms_gpu':>12}{'rms_ref':>12}")
    names = ["input_conv"] + [k for k in taps if k != "input_conv"]
    for name in names:
        ref = taps[name]
        try:
            got = eng.debug_activation(name, ref.size).cpu().numpy().reshape(ref.shape)
        except Exception as ex:  # noqa: BLE001
            print(f"{name:<18} ERROR {ex}")
            continue
        print(f"{name:<18}{util.rel_l2(got, ref):>12.3e}{util.max_rel(got, ref):>12.3e}"
              f"{float(np.sqrt((got.astype(np.float64)**2).mean())):>12.4f}{float(np.sqrt((ref.astype(np.float64)**2).mean())):>12.4f}")
    print(f"{'velocity':<18}{util.rel_l2(v, v_or):>12.3e}{util.max_rel(v, v_or):>12.3e}   (vs oracle bf16 policy)")
    print(f"{'velocity':<18}{util.rel_l2(v, g['v']):>12.3e}{util.max_rel(v, g['v']):>12.3e}   (vs reference fp32 golden)")
    print("nan count:", int(np.isnan(v).sum()), "launches:", eng.launch_count())


if __name__ == "__main__":
    main()
