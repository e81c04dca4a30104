"""CPU port of the hot path in functional PyTorch fp32.  TEST / BASELINE INFRASTRUCTURE ONLY.

Why a second restatement next to ``unet_oracle.py`` (numpy): the reference's arithmetic lives in PyTorch's CPU
kernels (oneDNN convolutions, ATen GroupNorm/softmax).  For the *reported CPU baseline* (bench.py ``cpu_baseline``
and ``--impl reference``) the honest comparison is the same library on the same host cores, so this file restates
the reference's call sequence with ``torch.nn.functional`` on a plain dict of tensors -- no nn.Module, no code
from /root/reference.  The reference itself cannot travel to the GPU box (it is a Python checkout that the rules
keep out of this repo); this port is pinned against its outputs in ``tests/test_oracle_golden.py``.

Only ``tests/`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs import this.

Citations (into /root/reference): models/unet.py:20-27 (sinusoid), :55-64 (ResidualBlock), :79-100 (attention),
:229-275 (UNet.forward); models/base_flow.py:158-173 (Euler loop).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def _res(P, pre, x, temb_act):
    h = F.conv2d(F.silu(F.group_norm(x, 8, P[pre + "norm1.weight"], P[pre + "norm1.bias"])),
                 P[pre + "conv1.weight"], P[pre + "conv1.bias"], padding=1)
    h = h + F.linear(temb_act, P[pre + "time_mlp.1.weight"], P[pre + "time_mlp.1.bias"])[:, :, None, None]
    h = F.conv2d(F.silu(F.group_norm(h, 8, P[pre + "norm2.weight"], P[pre + "norm2.bias"])),
                 P[pre + "conv2.weight"], P[pre + "conv2.bias"], padding=1)
    if (pre + "shortcut.weight") in P:
        x = F.conv2d(x, P[pre + "shortcut.weight"], P[pre + "shortcut.bias"])
    return h + x


def _attn(P, pre, x, heads=4):
    b, c, hh, ww = x.shape
    qkv = F.conv2d(F.group_norm(x, 8, P[pre + "norm.weight"], P[pre + "norm.bias"]), P[pre + "qkv.weight"], P[pre + "qkv.bias"])
    q, k, v = (a.reshape(b, heads, c // heads, hh * ww) for a in qkv.chunk(3, dim=1))
    att = torch.softmax(torch.einsum("bhcn,bhcm->bhnm", q, k) * (c // heads) ** -0.5, dim=-1)
    o = torch.einsum("bhnm,bhcm->bhcn", att, v).reshape(b, c, hh, ww)
    return x + F.conv2d(o, P[pre + "proj.weight"], P[pre + "proj.bias"])


def unet_forward_grad(P: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, model_channels: int = 64,
                      channel_mult: Sequence[int] = (1, 2, 4), num_res_blocks: int = 2, pre: str = "velocity_net."):
    """The forward pass without the no_grad guard (oracle/train_oracle.py differentiates through it)."""
    half = model_channels // 2
    freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * freqs[None, :]
    e = torch.cat((arg.sin(), arg.cos()), dim=-1)
    e = F.linear(F.silu(F.linear(e, P[pre + "time_mlp.1.weight"], P[pre + "time_mlp.1.bias"])),
                 P[pre + "time_mlp.3.weight"], P[pre + "time_mlp.3.bias"])
    ta = F.silu(e)
    h = F.conv2d(x, P[pre + "input_conv.weight"], P[pre + "input_conv.bias"], padding=1)
    skips, bi, nlev = [], 0, len(channel_mult)
    for lv in range(nlev):
        for _ in range(num_res_blocks):
            h = _res(P, f"{pre}enc_blocks.{bi}.", h, ta)
            bi += 1
        skips.append(h)
        if lv < nlev - 1:
            h = F.conv2d(h, P[f"{pre}downsamples.{lv}.weight"], P[f"{pre}downsamples.{lv}.bias"], stride=2, padding=1)
    h = _res(P, pre + "mid_block1.", h, ta)
    h = _attn(P, pre + "mid_attn.", h)
    h = _res(P, pre + "mid_block2.", h, ta)
    bi = 0
    for li in range(nlev):
        h = torch.cat([h, skips.pop()], dim=1)
        for _ in range(num_res_blocks):
            h = _res(P, f"{pre}dec_blocks.{bi}.", h, ta)
            bi += 1
        if li < nlev - 1:
            h = F.conv2d(F.interpolate(h, scale_factor=2, mode="nearest"), P[f"{pre}upsamples.{li}.1.weight"],
                         P[f"{pre}upsamples.{li}.1.bias"], padding=1)
    h = F.silu(F.group_norm(h, 8, P[pre + "output_conv.0.weight"], P[pre + "output_conv.0.bias"]))
    return F.conv2d(h, P[pre + "output_conv.2.weight"], P[pre + "output_conv.2.bias"], padding=1)


@torch.no_grad()
def unet_forward(P, x, t, **kw):
    return unet_forward_grad(P, x, t, **kw)


@torch.no_grad()
def euler_sample(P, noise: torch.Tensor, num_steps: int, **arch) -> torch.Tensor:
    x = noise
    dt = 1.0 / num_steps
    for i in range(num_steps):
        t = torch.ones(x.shape[0]) * (i * dt)
        x = x + unet_forward(P, x, t, **arch) * dt
    return x
