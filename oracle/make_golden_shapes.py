"""Golden vectors from the UNMODIFIED reference for shapes outside the two main golden cases (run in the build container only).

    python oracle/make_golden_shapes.py     # writes tests/golden/shapes_<name>.npz + tests/golden/shapes_manifest.json

Same recipe as oracle/make_golden.py (imports /root/reference/models; seeded weights re-created by the package's
bit-identical constructor; sha256 of the state_dict recorded): a 16-pixel two-level network, a grayscale and a 4-channel one.
Recorded per case: velocity, 3-step Euler sample, rectified-flow loss on seeded (x0, x1, t) and every parameter's gradient
norm plus a strided sample of each gradient from the reference's own ``loss.backward()``.

TEST INFRASTRUCTURE: nothing in the product path imports this.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

SHAPES = {
    # name: (image_size, in_channels, model_channels, channel_mult, num_res_blocks)   -- keep in step with tests/test_gpu_shapes.py
    "16px_two_levels": (16, 3, 64, [1, 2], 1),
    "gray_32px": (32, 1, 64, [1, 2, 4], 1),
    "four_channel_32px": (32, 4, 64, [1, 2, 4], 1),
}
SEED, INPUT_SEED, BATCH, STRIDE = 1234, 99, 5, 241


def main():
    from oracle.make_golden import state_sha
    sys.path.insert(0, "/root/reference")
    import models as ref  # the reference, unmodified
    sys.path.pop(0)
    import rectified_flow_vision_b200 as mine

    torch.set_num_threads(os.cpu_count())
    manifest = {"torch": torch.__version__, "seed": SEED, "input_seed": INPUT_SEED, "batch": BATCH, "stride": STRIDE, "cases": {}}
    for name, (S, cin, mc, mult, nres) in SHAPES.items():
        kw = dict(image_size=S, in_channels=cin, model_channels=mc, channel_mult=mult, num_res_blocks=nres)
        torch.manual_seed(SEED)
        rm = ref.RectifiedFlowModel(device="cpu", **kw)
        torch.manual_seed(SEED)
        mm = mine.RectifiedFlowModel(device="cpu", **kw)
        rsd, msd = rm.state_dict(), mm.state_dict()
        assert list(rsd.keys()) == list(msd.keys())
        for k in rsd:
            assert torch.equal(rsd[k], msd[k]), f"init mismatch at {k}"
        g = torch.Generator().manual_seed(INPUT_SEED)
        x = torch.randn(BATCH, cin, S, S, generator=g)
        x1 = torch.randn(BATCH, cin, S, S, generator=g)
        t = torch.rand(BATCH, generator=g)
        rm.eval()   # dropout off: the parity configuration
        out = {}
        with torch.no_grad():
            out["v"] = rm.forward(x, t).numpy()
            out["sample_3"] = rm.sample(noise=x, num_steps=3).numpy()
        xt, target = rm.get_interpolation(x, x1, t)
        pred = rm.forward(xt, t)
        loss = torch.nn.functional.mse_loss(pred, target)
        rm.zero_grad()
        loss.backward()
        names, norms = [], []
        for k, p in rm.named_parameters():
            names.append(k)
            gflat = p.grad.detach().reshape(-1)
            norms.append(float(gflat.double().norm()))
            out["grad_sampled/" + k] = gflat[::STRIDE].numpy().copy()
        np.savez_compressed(os.path.join(GOLD, f"shapes_{name}.npz"), x=x.numpy(), x1=x1.numpy(), t=t.numpy(), loss=np.float64(loss.item()),
                            names=np.array(names), grad_norm_per_tensor=np.array(norms), **out)
        manifest["cases"][name] = {"kwargs": kw, "state_sha256": state_sha(rsd), "loss": float(loss.item())}
        print(name, "ok; loss", float(loss.item()), "params", sum(p.numel() for p in rm.parameters()))
    with open(os.path.join(GOLD, "shapes_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
