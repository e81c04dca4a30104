"""Autograd surface of the native training path: ``UNet.forward`` in training mode and ``loss.backward()``.

The reference trains through PyTorch autograd (``loss = compute_loss(x); optimizer.zero_grad(); loss.backward();
clip_grad_norm_(model.parameters(), 1.0); optimizer.step()``, models/base_flow.py:266-275, models/rectified_flow.py:217-238).
The fast path of this package (``training.NativeTrainer``) never builds a graph, but the contract of SURVEY §8(b) also
lists ``parameters()`` / ``.grad`` / ``torch.optim.AdamW``.  This module closes it: the training-mode velocity is a
``torch.autograd.Function`` whose forward is ``rfv_train_forward`` (dropout active, activations of the micro-batch kept
on the device) and whose backward is ``rfv_train_backward`` (dL/dv in, the hand-written backward pass, dL/dparam out of the
engine's flat gradient buffer into the 174 ``Parameter.grad`` tensors).  Nothing here computes with PyTorch kernels: torch
only carries the loss the caller writes on top of ``v`` (e.g. ``F.mse_loss``) and the optimizer the caller chose.

Batches larger than the engine's micro-batch, and backward calls that arrive after another forward on the same engine,
recompute the forward of each micro-batch (same dropout seed) right before its backward: activations are kept for one
micro-batch only.
"""
from __future__ import annotations

from typing import List

import torch


def _draw_seed() -> int:
    """Dropout seed from torch's global CPU generator: ``torch.manual_seed`` makes training-mode forwards reproducible,
    as it does for nn.Dropout in the reference."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class _VelocityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unet, x, t, dropout_p, seed, *params):
        eng = unet.train_engine(x.shape[-1], x.device)
        x = eng._images(x.detach(), "x")
        t = eng._times(t.detach(), x.shape[0])
        v = torch.empty_like(x)
        mb = eng.micro_batch
        bounds = [(b0, min(b0 + mb, x.shape[0])) for b0 in range(0, x.shape[0], mb)]
        for i, (b0, b1) in enumerate(bounds):
            eng.train_forward(x[b0:b1], t[b0:b1], dropout_p, seed + i, v[b0:b1])
        ctx.unet, ctx.eng, ctx.x, ctx.t = unet, eng, x, t
        ctx.dropout_p, ctx.seed, ctx.bounds = dropout_p, seed, bounds
        ctx.token = eng.fwd_token
        ctx.names = [n for n, _ in unet.named_parameters()]
        return v

    @staticmethod
    def backward(ctx, dv):
        eng, x, t = ctx.eng, ctx.x, ctx.t
        dv = dv.contiguous().float()
        eng.zero_grad()
        kept = len(ctx.bounds) == 1 and eng.fwd_token == ctx.token
        scratch = None
        for i, (b0, b1) in enumerate(ctx.bounds):
            if not kept:
                if scratch is None:
                    scratch = torch.empty_like(x[b0:b1])
                eng.train_forward(x[b0:b1], t[b0:b1], ctx.dropout_p, ctx.seed + i, scratch[: b1 - b0])
            eng.train_backward(dv[b0:b1])
        numel = dict(eng.tensor_names)
        grads: List[torch.Tensor] = []
        params = dict(ctx.unet.named_parameters())
        for i, name in enumerate(ctx.names):
            if not ctx.needs_input_grad[5 + i]:
                grads.append(None)
                continue
            full = "velocity_net." + name
            grads.append(eng.get_grad(full, numel[full]).view_as(params[name]))
        return (None, None, None, None, None, *grads)


def velocity_with_grad(unet, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """``unet(x, t)`` in training mode: dropout on, differentiable with respect to every parameter of ``unet``."""
    if x.dim() != 4 or x.shape[1] != unet.in_channels or x.shape[2] != x.shape[3]:
        raise ValueError(f"expected x of shape [B,{unet.in_channels},S,S], got {tuple(x.shape)}")
    if t.dim() != 1 or t.shape[0] != x.shape[0]:
        raise ValueError(f"expected t of shape [{x.shape[0]}], got {tuple(t.shape)}")
    if x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("the native backward pass produces parameter gradients only: x must not require grad "
                                  "(no call site of the reference differentiates with respect to the input)")
    params = [p for _, p in unet.named_parameters()]
    return _VelocityFn.apply(unet, x, t, float(unet.dropout_p), _draw_seed(), *params)
