"""CPU oracle of the reflow / flow-matching TRAINING STEP.  TEST INFRASTRUCTURE ONLY (tests/, bench.py cpu legs).

Restates, on the functional fp32 port of the network (oracle/torch_port.py), what one iteration of the reference
training loops computes (models/rectified_flow.py:217-238, models/base_flow.py:113-129,266-275):

    x_t = (1-t) x0 + t x1;  target = x1 - x0                 (models/base_flow.py:81-89)
    loss = mean((v(x_t, t) - target)^2)                       (models/rectified_flow.py:231)
    loss.backward()                                           (:235)   -> torch.autograd over the functional port
    clip_grad_norm_(params, 1.0)                              (:236)   -> restated below
    AdamW(lr, betas (0.9, 0.999), eps 1e-8, weight_decay 0.01).step()   (:208, :237)   -> restated below

Dropout is the identity here (p = 0): the reference's nn.Dropout draws from torch's generator, which no other
implementation can reproduce bit-for-bit; parity of the training step is therefore pinned with dropout disabled
(SURVEY §8d config 4 "dropout 0.0 for parity runs").  Pinned against the unmodified reference's own
loss / gradients / post-step parameters by oracle/make_golden_train.py -> tests/golden/train_*.npz.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import torch_port


def loss_and_grads(P: Dict[str, torch.Tensor], x0: torch.Tensor, x1: torch.Tensor, t: torch.Tensor,
                   **arch) -> Tuple[float, Dict[str, torch.Tensor]]:
    Q = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    tt = t.view(-1, 1, 1, 1)
    x_t = (1 - tt) * x0 + tt * x1
    target = x1 - x0
    pred = torch_port.unet_forward_grad(Q, x_t, t, **arch)
    loss = torch.mean((pred - target) ** 2)
    grads = torch.autograd.grad(loss, list(Q.values()))
    return float(loss.item()), dict(zip(Q.keys(), grads))


def clip_coef(grads: Dict[str, torch.Tensor], max_norm: float = 1.0) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_: total 2-norm over all gradients, coefficient clamp(max_norm/(norm+1e-6), 1)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_step(P: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict[str, Dict[str, torch.Tensor]],
               step: int, lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
               weight_decay: float = 0.01, max_norm: float = 1.0) -> float:
    """In-place clip + AdamW update of P (decoupled weight decay, bias-corrected moments); returns the pre-clip norm."""
    total, coef = clip_coef(grads, max_norm)
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    for k, p in P.items():
        g = grads[k] * coef
        st = state.setdefault(k, {"m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["m"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["v"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        p.mul_(1 - lr * weight_decay)
        denom = (st["v"] / bc2).sqrt_().add_(eps)
        p.addcdiv_(st["m"] / bc1, denom, value=-lr)
    return total
