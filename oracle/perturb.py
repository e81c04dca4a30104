"""Seeded perturbation of a checkpoint so that parity is pinned on NON-TRIVIAL weights.  TEST INFRASTRUCTURE ONLY.

Why: the reference constructor (models/unet.py:35-52,75-77,223-227) leaves every GroupNorm affine at gamma = 1, beta = 0
(torch's default) and every conv / linear bias tiny (uniform in +-1/sqrt(fan_in)).  The checkpoints the reference's numbers
are quoted on (`.MISSING_LARGE_BLOBS`: base_flow_final.pt, rectified_flow_k1_final.pt, ...) are absent from the checkout
and 45 MB each, so they cannot be committed either.  A wrong gamma / beta channel offset -- across the decoder's virtual
concat, in the fused GroupNorm coefficient table, in the backward pass -- is invisible with gamma = 1, beta = 0.

`perturb_state_dict` rewrites a state_dict deterministically from a CPU torch.Generator (bit-reproducible on any box with
the same torch): every GroupNorm weight becomes 1 + 0.5 N(0,1), every GroupNorm bias 0.5 N(0,1), every other bias gets
0.2 N(0,1) added, conv / linear weights keep their seeded initialisation.  The goldens in tests/golden/pert_*.npz are the
UNMODIFIED reference's outputs after `load_state_dict(perturbed)` (oracle/make_golden_weights.py); the GPU tests apply the
same function to the package model.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

GN_WEIGHT_STD, GN_BIAS_STD, BIAS_STD = 0.5, 0.5, 0.2


def is_group_norm(key: str) -> bool:
    """GroupNorm sites of the UNet: ResidualBlock.norm1/norm2, AttentionBlock.norm, output_conv.0
    (models/unet.py:36,40,75,224)."""
    return ".norm1." in key or ".norm2." in key or ".norm." in key or ".output_conv.0." in key


def perturb_state_dict(sd, seed: int = 123):
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        v = v.detach().clone().float().cpu()
        if is_group_norm(k):
            r = torch.randn(v.shape, generator=g)
            v = (1.0 + GN_WEIGHT_STD * r) if k.endswith(".weight") else GN_BIAS_STD * r
        elif k.endswith(".bias"):
            v = v + BIAS_STD * torch.randn(v.shape, generator=g)
        out[k] = v
    return out
