"""Golden vectors of the TRAINING STEP from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden_train.py      # writes tests/golden/train_<case>.npz

For the seeded weights / inputs of oracle/make_golden.py it runs the reference's own training-step body
(models/rectified_flow.py:222-237: get_interpolation, forward, F.mse_loss, backward, clip_grad_norm_(1.0),
torch.optim.AdamW.step) for three consecutive steps on a fixed batch, dropout disabled (model.eval(): GroupNorm has
no train/eval difference, nn.Dropout becomes the identity), and records the loss of every step, the pre-clip gradient
norm, the per-tensor gradient norms of step 1 (174 values), a handful of complete small gradients, and per-tensor
parameter norms plus the full update of selected tensors after the three steps.  LR is 1e-3 (10x the reference's
default) so that three steps move the parameters measurably.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import CASES, GOLD, seeded_inputs  # noqa: E402

FULL = ["velocity_net.output_conv.2.weight", "velocity_net.output_conv.2.bias", "velocity_net.input_conv.weight",
        "velocity_net.input_conv.bias", "velocity_net.time_mlp.1.bias", "velocity_net.time_mlp.3.bias",
        "velocity_net.enc_blocks.0.norm1.weight", "velocity_net.enc_blocks.0.norm2.bias",
        "velocity_net.enc_blocks.0.time_mlp.1.bias", "velocity_net.enc_blocks.0.conv1.bias",
        "velocity_net.mid_attn.norm.weight", "velocity_net.mid_attn.qkv.bias", "velocity_net.mid_attn.proj.bias",
        "velocity_net.dec_blocks.0.shortcut.bias", "velocity_net.dec_blocks.0.norm1.weight",
        "velocity_net.downsamples.0.bias", "velocity_net.upsamples.0.1.bias", "velocity_net.output_conv.0.weight"]
SAMPLED = ["velocity_net.enc_blocks.0.conv1.weight", "velocity_net.enc_blocks.0.conv2.weight",
           "velocity_net.downsamples.0.weight", "velocity_net.upsamples.0.1.weight", "velocity_net.mid_attn.qkv.weight",
           "velocity_net.mid_attn.proj.weight", "velocity_net.dec_blocks.0.conv1.weight",
           "velocity_net.dec_blocks.0.shortcut.weight", "velocity_net.mid_block1.conv2.weight",
           "velocity_net.enc_blocks.0.time_mlp.1.weight", "velocity_net.time_mlp.3.weight"]
LR, STEPS, STRIDE = 1e-3, 3, 61


def main():
    sys.path.insert(0, "/root/reference")
    import models as ref
    sys.path.pop(0)
    torch.set_num_threads(os.cpu_count())
    for name, (kw, batch, seed) in CASES.items():
        torch.manual_seed(seed)
        rm = ref.RectifiedFlowModel(device="cpu", **kw)
        rm.eval()  # dropout off; nothing else depends on the mode
        x0, t, x1 = seeded_inputs(batch, rm.in_channels, rm.image_size)
        names = [k for k, _ in rm.named_parameters()]
        opt = torch.optim.AdamW(rm.parameters(), lr=LR)
        out = {"losses": [], "grad_norms_total": []}
        p0 = {k: v.detach().clone() for k, v in rm.named_parameters()}
        for step in range(STEPS):
            x_t, target = rm.get_interpolation(x0, x1, t)
            pred = rm.forward(x_t, t)
            loss = torch.nn.functional.mse_loss(pred, target)
            opt.zero_grad()
            loss.backward()
            if step == 0:
                out["grad_norm_per_tensor"] = np.array([float(p.grad.norm()) for _, p in rm.named_parameters()])
                for k, p in rm.named_parameters():
                    if k in FULL:
                        out["grad_full/" + k] = p.grad.numpy().copy()
                    if k in SAMPLED:
                        out["grad_sampled/" + k] = p.grad.numpy().reshape(-1)[::STRIDE].copy()
            total = torch.nn.utils.clip_grad_norm_(rm.parameters(), 1.0)
            opt.step()
            out["losses"].append(float(loss.item()))
            out["grad_norms_total"].append(float(total))
        out["param_norm_after"] = np.array([float(p.detach().norm()) for _, p in rm.named_parameters()])
        out["update_norm"] = np.array([float((p.detach() - p0[k]).norm()) for k, p in rm.named_parameters()])
        for k, p in rm.named_parameters():
            if k in FULL:
                out["update_full/" + k] = (p.detach() - p0[k]).numpy().copy()
        out["losses"] = np.array(out["losses"])
        out["grad_norms_total"] = np.array(out["grad_norms_total"])
        np.savez_compressed(os.path.join(GOLD, f"train_{name}.npz"), names=np.array(names), lr=LR, stride=STRIDE, **out)
        print(name, "losses", out["losses"], "grad norms", out["grad_norms_total"])


if __name__ == "__main__":
    main()
