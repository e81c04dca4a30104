"""Golden vectors on NON-TRIVIAL weights, long integrations and 128x128, from the UNMODIFIED reference (build container only).

    python oracle/make_golden_weights.py     # writes tests/golden/pert_*.npz, default64_100.npz, weights_manifest.json

What the first golden set (oracle/make_golden.py) cannot see: it uses constructor-initialised weights, where every GroupNorm
has gamma = 1, beta = 0 and all biases are tiny, it stops at 8 Euler steps, and it has no 128x128 case.  Here the reference
model is built with the same seed, its state_dict is rewritten by oracle/perturb.py (seeded: GroupNorm gamma ~ 1 + 0.5 N,
beta ~ 0.5 N, every other bias + 0.2 N) and loaded back through the reference's own `load_state_dict`; then the reference's
own code produces
  * the velocity and per-layer fingerprints                      models/base_flow.py:91-102, models/unet.py:229-275
  * 1-, 8- and 100-step Euler samples (100 = the step count of pair generation, experiments/train_rectified.py:76-80)
                                                                  models/base_flow.py:133-177
  * the flow-matching loss, straightness                          models/base_flow.py:81-89, models/rectified_flow.py:82-124
  * the gradients of one training-step body (loss.backward())     models/rectified_flow.py:222-233
for the default 64x64 net, the small 32x32 net and the default architecture at 128x128 (BASELINE.json configs[4]).
`default64_100.npz` adds the 100-step sample of the UNperturbed default64 case.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, seeded_inputs, state_sha  # noqa: E402
from oracle.make_golden_train import FULL, SAMPLED, STRIDE  # noqa: E402
from oracle.perturb import perturb_state_dict  # noqa: E402

PERT_SEED = 123
CASES = {
    # name: (ctor kwargs, batch, weight seed, Euler step counts)
    "pert_default64": (dict(image_size=64), 2, 0, (1, 8, 100)),
    "pert_small32": (dict(image_size=32, model_channels=64, channel_mult=[1, 2], num_res_blocks=1), 3, 1, (1, 8, 100)),
    "pert_default128": (dict(image_size=128), 2, 2, (1, 8)),
}


def hooked_modules(net):
    hooked = {"input_conv": net.input_conv, "mid_block1": net.mid_block1, "mid_attn": net.mid_attn, "mid_block2": net.mid_block2}
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
    return hooked


def main():
    sys.path.insert(0, "/root/reference")
    import models as ref  # the reference, unmodified
    sys.path.pop(0)
    torch.set_num_threads(os.cpu_count())
    manifest = {"torch": torch.__version__, "perturb_seed": PERT_SEED, "cases": {}}
    for name, (kw, batch, seed, step_list) in CASES.items():
        torch.manual_seed(seed)
        rm = ref.RectifiedFlowModel(device="cpu", **kw)
        init_sha = state_sha(rm.state_dict())
        rm.load_state_dict(perturb_state_dict(rm.state_dict(), PERT_SEED))
        rm.eval()
        sd = rm.state_dict()
        c, s = rm.in_channels, rm.image_size
        x, t, x1 = seeded_inputs(batch, c, s)
        out, taps, hooks = {}, {}, []
        for n_, m_ in hooked_modules(rm.velocity_net).items():
            hooks.append(m_.register_forward_hook(lambda mod, inp, o, n_=n_: taps.__setitem__(n_, o.detach().clone())))
        with torch.no_grad():
            out["v"] = rm.forward(x, t).numpy()
        for h in hooks:
            h.remove()
        tapinfo = {}
        for n_, o in taps.items():
            a = o.numpy().astype(np.float64)
            tapinfo[n_] = {"shape": list(a.shape), "mean": float(a.mean()), "rms": float(np.sqrt((a ** 2).mean()))}
            out["tap_" + n_] = o.numpy().reshape(-1)[::997].copy()
        with torch.no_grad():
            for steps in step_list:
                out[f"sample_{steps}"] = rm.sample(noise=x, num_steps=steps).numpy()
            straight = float(rm.compute_straightness(x, x1, num_points=3))
        # one training-step body, dropout off (eval mode), exactly as oracle/make_golden_train.py does for the seeded weights
        x_t, target = rm.get_interpolation(x, x1, t)
        pred = rm.forward(x_t, t)
        loss = torch.nn.functional.mse_loss(pred, target)
        rm.zero_grad()
        loss.backward()
        names = [k for k, _ in rm.named_parameters()]
        out["grad_norm_per_tensor"] = np.array([float(p.grad.norm()) for _, p in rm.named_parameters()])
        for k, p in rm.named_parameters():
            if k in FULL or ".norm" in k or "output_conv.0" in k:   # every GroupNorm gradient in full: they are small
                out["grad_full/" + k] = p.grad.numpy().copy()
            if k in SAMPLED:
                out["grad_sampled/" + k] = p.grad.numpy().reshape(-1)[::STRIDE].copy()
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), x=x.numpy(), t=t.numpy(), x1=x1.numpy(), names=np.array(names),
                            stride=STRIDE, **out)
        manifest["cases"][name] = {"kwargs": kw, "batch": batch, "seed": seed, "input_seed": 42, "steps": list(step_list),
                                   "init_sha256": init_sha, "state_sha256": state_sha(sd), "fm_loss": float(loss.item()),
                                   "straightness_3": straight, "taps": tapinfo}
        print(name, "ok; loss", float(loss.item()), "straightness", straight, "|v| rms", float(np.sqrt((out["v"] ** 2).mean())))
    # 100-step sample of the unperturbed default64 case (its inputs live in default64.npz)
    torch.manual_seed(0)
    rm = ref.BaseFlowModel(device="cpu", image_size=64)
    x, _, _ = seeded_inputs(2, 3, 64)
    with torch.no_grad():
        s100 = rm.sample(noise=x, num_steps=100).numpy()
    np.savez_compressed(os.path.join(GOLD, "default64_100.npz"), sample_100=s100)
    manifest["default64_100"] = {"state_sha256": state_sha(rm.state_dict())}
    with open(os.path.join(GOLD, "weights_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
