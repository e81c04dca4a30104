"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz + tests/golden/manifest.json

Imports /root/reference/models (needs only torch / numpy / tqdm) and records, for seeded weights and seeded
host noise, the reference's own outputs on the hot path.  The GPU box has no /root/reference, so the vectors are
committed; weights are NOT stored (45 MB): they are re-created by ``torch.manual_seed(seed)`` + the package's
parameter container, whose initialisation is bit-identical to the reference constructor's -- this script asserts
that and stores a sha256 of the state_dict so the tests can re-check it anywhere.

TEST INFRASTRUCTURE: nothing in the product path imports this.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (ctor kwargs, batch, seed)
    "default64": (dict(image_size=64), 2, 0),
    "small32": (dict(image_size=32, model_channels=64, channel_mult=[1, 2], num_res_blocks=1), 3, 1),
}


def state_sha(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def seeded_inputs(batch, c, s, seed=42):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, c, s, s, generator=g)
    t = torch.rand(batch, generator=g)
    x1 = torch.randn(batch, c, s, s, generator=g)
    return x, t, x1


def main():
    sys.path.insert(0, "/root/reference")
    import models as ref  # the reference, unmodified
    sys.path.pop(0)
    import rectified_flow_vision_b200 as mine

    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"torch": torch.__version__, "cases": {}}
    for name, (kw, batch, seed) in CASES.items():
        torch.manual_seed(seed)
        rm = ref.BaseFlowModel(device="cpu", **kw)
        torch.manual_seed(seed)
        mm = mine.BaseFlowModel(device="cpu", **kw)
        rsd, msd = rm.state_dict(), mm.state_dict()
        assert list(rsd.keys()) == list(msd.keys()), "state_dict key order differs"
        for k in rsd:
            assert rsd[k].shape == msd[k].shape and torch.equal(rsd[k], msd[k]), f"init mismatch at {k}"
        rm.eval()
        c, s = rm.in_channels, rm.image_size
        x, t, x1 = seeded_inputs(batch, c, s)
        out = {}
        taps = {}
        hooks = []
        net = rm.velocity_net
        hooked = {"input_conv": net.input_conv, "mid_block1": net.mid_block1, "mid_attn": net.mid_attn,
                  "mid_block2": net.mid_block2}
        for i, b in enumerate(net.enc_blocks):
            hooked[f"enc_blocks.{i}"] = b
        for i, b in enumerate(net.dec_blocks):
            hooked[f"dec_blocks.{i}"] = b
        for i, b in enumerate(net.downsamples):
            if b is not None:
                hooked[f"downsamples.{i}"] = b
        for i, b in enumerate(net.upsamples):
            if b is not None:
                hooked[f"upsamples.{i}"] = b
        for n_, m_ in hooked.items():
            hooks.append(m_.register_forward_hook(lambda mod, inp, o, n_=n_: taps.__setitem__(n_, o.detach().clone())))
        with torch.no_grad():
            out["v"] = rm.forward(x, t).numpy()
        for h in hooks:
            h.remove()
        # per-layer fingerprints (full tensors would be tens of MB): mean, rms, and a strided sample
        tapinfo = {}
        for n_, o in taps.items():
            a = o.numpy().astype(np.float64)
            tapinfo[n_] = {"shape": list(a.shape), "mean": float(a.mean()), "rms": float(np.sqrt((a ** 2).mean()))}
            out["tap_" + n_] = o.numpy().reshape(-1)[::997].copy()
        with torch.no_grad():
            for steps in (1, 2, 4, 8):
                out[f"sample_{steps}"] = rm.sample(noise=x, num_steps=steps).numpy()
            traj = rm.sample_with_trajectory(x, num_steps=4, save_every=2)
            out["traj_4_2"] = np.stack([a.numpy() for a in traj])
            xt, target = rm.get_interpolation(x, x1, t)
            out["xt"], out["target"] = xt.numpy(), target.numpy()
            pred = rm.forward(xt, t)
            loss = torch.nn.functional.mse_loss(pred, target).item()
        rr = ref.RectifiedFlowModel(device="cpu", **kw)
        rr.load_state_dict(rsd)
        straight = float(rr.compute_straightness(x, x1, num_points=3))
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), x=x.numpy(), t=t.numpy(), x1=x1.numpy(), **out)
        manifest["cases"][name] = {"kwargs": kw, "batch": batch, "seed": seed, "input_seed": 42,
                                   "state_sha256": state_sha(rsd), "num_params": ref.count_parameters(rm),
                                   "fm_loss": loss, "straightness_3": straight, "taps": tapinfo}
        print(name, "ok; params", ref.count_parameters(rm), "loss", loss, "straightness", straight)
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
