"""CPU oracle (numpy, fp32) for the Euler-integration hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (``rectified_flow_vision_b200``) never does; it fails
loudly when the CUDA library is missing.

What this restates (all citations are into /root/reference):
  * sinusoidal time embedding                   models/unet.py:20-27
  * time MLP (Linear-SiLU-Linear)               models/unet.py:157-162
  * ResidualBlock                               models/unet.py:55-64
  * AttentionBlock                              models/unet.py:79-100
  * UNet wiring (encoder / middle / decoder)    models/unet.py:229-275
  * Euler sampler                               models/base_flow.py:158-177
  * trajectory sampler                          models/base_flow.py:196-208
  * linear interpolation + target               models/base_flow.py:81-89
  * flow-matching MSE loss                      models/base_flow.py:129, models/rectified_flow.py:231
  * straightness metric                         models/rectified_flow.py:98-124
  * reflow pair generation batching             models/rectified_flow.py:150-170

The arithmetic of the reference lives in third-party PyTorch (requirements.txt:1 pins torch==2.1.0; the
image has 2.11.0): Conv2d = cross-correlation + bias, GroupNorm = biased variance with eps 1e-5, SiLU =
x*sigmoid(x), softmax over keys, nearest-neighbour x2 upsampling.  Those published definitions are what is
restated below in numpy.

Parity pin: the reference ships no golden vectors for this path, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by ``oracle/make_golden.py``
(imports /root/reference/models unmodified) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them.

``policy`` argument: ``None`` computes everything in fp32 (the reference's arithmetic).  ``BF16_POLICY``
additionally rounds to bfloat16 at exactly the points where the CUDA path stores bf16 (weights of the
tensor-core convs, every activation written to HBM, the softmax probabilities), so a kernel bug is not hidden
inside the bf16 tolerance.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------
# precision policies
# ----------------------------------------------------------------------------------------------------------
def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what ``__float2bfloat16_rn`` does on the device)."""
    a = np.ascontiguousarray(a, dtype=F32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(F32).reshape(a.shape)
    return np.where(np.isfinite(a), out, a).astype(F32)


class Policy:
    """Where values get rounded.  ``act``: activations stored to HBM; ``w``: tensor-core conv weights;
    ``p``: softmax probabilities before the P.V product."""

    def __init__(self, name: str, rounder=None):
        self.name = name
        self._r = rounder

    def act(self, a):
        return a if self._r is None else self._r(a)

    w = act
    p = act


FP32_POLICY = Policy("fp32")
BF16_POLICY = Policy("bf16", round_bf16)


# ----------------------------------------------------------------------------------------------------------
# elementary ops
# ----------------------------------------------------------------------------------------------------------
def silu(x: np.ndarray) -> np.ndarray:
    return (x / (1.0 + np.exp(-x.astype(F32)))).astype(F32)


def linear(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.Linear: y = x @ w.T + b."""
    return (x.astype(F32) @ w.astype(F32).T + b.astype(F32)).astype(F32)


def sinusoidal_embedding(t: np.ndarray, dim: int) -> np.ndarray:
    """models/unet.py:20-27.  t is NOT scaled by 1000; frequencies exp(-j*ln(1e4)/(half-1))."""
    half = dim // 2
    step = F32(math.log(10000) / (half - 1))
    freqs = np.exp(np.arange(half, dtype=F32) * -step).astype(F32)
    arg = t.astype(F32)[:, None] * freqs[None, :]
    return np.concatenate([np.sin(arg), np.cos(arg)], axis=-1).astype(F32)


def group_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, groups: int = 8, eps: float = 1e-5,
               stats_from: Optional[np.ndarray] = None) -> np.ndarray:
    """nn.GroupNorm(groups, C): per (n, group) mean / biased variance over (C/groups, H, W).

    ``stats_from`` lets the statistics come from a different (un-rounded) copy of the tensor than the one
    being normalised, which is what the fused CUDA path does (stats from the fp32 accumulators, normalise the
    bf16 tensor)."""
    n, c, h, w = x.shape
    src = x if stats_from is None else stats_from
    g = src.reshape(n, groups, -1).astype(np.float64)
    mean = g.mean(axis=2)
    var = g.var(axis=2)
    rstd = 1.0 / np.sqrt(var + eps)
    xg = x.reshape(n, groups, -1).astype(np.float64)
    y = ((xg - mean[:, :, None]) * rstd[:, :, None]).reshape(n, c, h, w)
    y = y * gamma.astype(np.float64)[None, :, None, None] + beta.astype(np.float64)[None, :, None, None]
    return y.astype(F32)


def conv2d(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int = 1, pad: int = 0) -> np.ndarray:
    """nn.Conv2d (cross-correlation, zero padding) as im2col + one GEMM.  x NCHW, w OIHW."""
    n, c, h, wd = x.shape
    o, ci, kh, kw = w.shape
    assert ci == c
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.zeros((n, c, h + 2 * pad, wd + 2 * pad), dtype=F32)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    cols = np.empty((n, c, kh, kw, ho, wo), dtype=F32)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, i, j] = xp[:, :, i:i + stride * ho:stride, j:j + stride * wo:stride]
    cols = cols.reshape(n, c * kh * kw, ho * wo)
    y = np.matmul(w.reshape(o, -1).astype(F32)[None], cols)  # [n, o, ho*wo]
    if b is not None:
        y = y + b.astype(F32)[None, :, None]
    return y.reshape(n, o, ho, wo).astype(F32)


def upsample_nearest2x(x: np.ndarray) -> np.ndarray:
    """nn.Upsample(scale_factor=2, mode='nearest'): out[y, x] = in[y // 2, x // 2]."""
    return x.repeat(2, axis=2).repeat(2, axis=3)


# ----------------------------------------------------------------------------------------------------------
# network blocks
# ----------------------------------------------------------------------------------------------------------
class UNetSpec:
    """Architecture hyper-parameters (models/unet.py:136-145).  attention_resolutions is accepted and
    ignored exactly like the reference (only mid_attn exists, models/unet.py:192)."""

    def __init__(self, in_channels=3, model_channels=64, out_channels=3, channel_mult=(1, 2, 4),
                 num_res_blocks=2, num_heads=4, groups=8):
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.channel_mult = tuple(channel_mult)
        self.num_res_blocks = num_res_blocks
        self.num_heads = num_heads
        self.groups = groups
        self.level_channels = [model_channels * m for m in self.channel_mult]
        self.time_dim = model_channels * 4


def _res_block(P: Dict[str, np.ndarray], pre: str, x: np.ndarray, temb_act: np.ndarray, pol: Policy,
               x_stats: Optional[np.ndarray]) -> (np.ndarray, np.ndarray):
    """models/unet.py:55-64 in eval mode (dropout = identity).  Returns (rounded output, un-rounded output);
    the un-rounded copy feeds the next GroupNorm's statistics under the bf16 policy."""
    a1 = pol.act(silu(group_norm(x, P[pre + "norm1.weight"], P[pre + "norm1.bias"], stats_from=x_stats)))
    h = conv2d(a1, pol.w(P[pre + "conv1.weight"]), P[pre + "conv1.bias"], 1, 1)
    tp = linear(temb_act, P[pre + "time_mlp.1.weight"], P[pre + "time_mlp.1.bias"])
    h_full = h + tp[:, :, None, None]
    h = pol.act(h_full)
    a2 = pol.act(silu(group_norm(h, P[pre + "norm2.weight"], P[pre + "norm2.bias"], stats_from=h_full)))
    out = conv2d(a2, pol.w(P[pre + "conv2.weight"]), P[pre + "conv2.bias"], 1, 1)
    if (pre + "shortcut.weight") in P:
        out = out + conv2d(x, pol.w(P[pre + "shortcut.weight"]), P[pre + "shortcut.bias"], 1, 0)
    else:
        out = out + x
    return pol.act(out), out


def _attention(P, pre: str, x: np.ndarray, heads: int, pol: Policy, x_stats) -> (np.ndarray, np.ndarray):
    """models/unet.py:79-100.  GroupNorm WITHOUT SiLU; head = contiguous C/heads channel block after a 3-way
    channel chunk; scores scaled by d^-0.5; softmax over keys."""
    n, c, h, w = x.shape
    d = c // heads
    hn = pol.act(group_norm(x, P[pre + "norm.weight"], P[pre + "norm.bias"], stats_from=x_stats))
    qkv = pol.act(conv2d(hn, pol.w(P[pre + "qkv.weight"]), P[pre + "qkv.bias"], 1, 0))
    q, k, v = [a.reshape(n, heads, d, h * w) for a in np.split(qkv, 3, axis=1)]
    s = np.einsum("bhcn,bhcm->bhnm", q, k).astype(F32) * F32(d ** -0.5)
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    den = e.sum(axis=-1, keepdims=True)
    if pol is FP32_POLICY:
        o = np.einsum("bhnm,bhcm->bhcn", (e / den).astype(F32), v)
    else:
        # flash-style: un-normalised probabilities are rounded, the row sum stays fp32
        o = np.einsum("bhnm,bhcm->bhcn", pol.p(e.astype(F32)), v) / np.swapaxes(den, 2, 3)
    o = pol.act(o.reshape(n, c, h, w).astype(F32))
    out = x + conv2d(o, pol.w(P[pre + "proj.weight"]), P[pre + "proj.bias"], 1, 0)
    return pol.act(out), out


def time_embedding(P, t: np.ndarray, spec: UNetSpec, pre: str = "velocity_net.") -> np.ndarray:
    """models/unet.py:157-162,231 -> [B, 4*model_channels], always fp32."""
    e = sinusoidal_embedding(t, spec.model_channels)
    e = linear(e, P[pre + "time_mlp.1.weight"], P[pre + "time_mlp.1.bias"])
    return linear(silu(e), P[pre + "time_mlp.3.weight"], P[pre + "time_mlp.3.bias"])


def unet_forward(P: Dict[str, np.ndarray], x: np.ndarray, t: np.ndarray, spec: Optional[UNetSpec] = None,
                 policy: Policy = FP32_POLICY, pre: str = "velocity_net.",
                 taps: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
    """UNet.forward, models/unet.py:229-275.  ``P`` is the reference state_dict as numpy arrays.
    ``taps`` (optional dict) receives named intermediate activations for per-layer tests."""
    spec = spec or UNetSpec()
    pol = policy
    x = np.asarray(x, dtype=F32)
    t = np.asarray(t, dtype=F32)
    temb_act = silu(time_embedding(P, t, spec, pre))  # every block applies SiLU first (unet.py:43-46)

    def tap(name, val):
        if taps is not None:
            taps[name] = val

    # the input conv runs on the tensor cores too: x and its weights are rounded to bf16 on the way into the MMA
    h_full = conv2d(pol.act(x), pol.w(P[pre + "input_conv.weight"]), P[pre + "input_conv.bias"], 1, 1)
    h = pol.act(h_full)
    tap("input_conv", h)

    skips: List = []
    bi = 0
    nlev = len(spec.channel_mult)
    for level in range(nlev):
        for _ in range(spec.num_res_blocks):
            h, h_full = _res_block(P, f"{pre}enc_blocks.{bi}.", h, temb_act, pol, h_full)
            tap(f"enc_blocks.{bi}", h)
            bi += 1
        skips.append((h, h_full))  # pushed BEFORE the downsample (unet.py:245)
        if level < nlev - 1:
            h_full = conv2d(h, pol.w(P[f"{pre}downsamples.{level}.weight"]), P[f"{pre}downsamples.{level}.bias"], 2, 1)
            h = pol.act(h_full)
            tap(f"downsamples.{level}", h)

    h, h_full = _res_block(P, pre + "mid_block1.", h, temb_act, pol, h_full)
    tap("mid_block1", h)
    h, h_full = _attention(P, pre + "mid_attn.", h, spec.num_heads, pol, h_full)
    tap("mid_attn", h)
    h, h_full = _res_block(P, pre + "mid_block2.", h, temb_act, pol, h_full)
    tap("mid_block2", h)

    bi = 0
    for li in range(nlev):
        sk, sk_full = skips.pop()
        h = np.concatenate([h, sk], axis=1)  # h first, skip second (unet.py:262)
        h_full = np.concatenate([h_full, sk_full], axis=1)
        for _ in range(spec.num_res_blocks):
            h, h_full = _res_block(P, f"{pre}dec_blocks.{bi}.", h, temb_act, pol, h_full)
            tap(f"dec_blocks.{bi}", h)
            bi += 1
        if li < nlev - 1:
            h_full = conv2d(upsample_nearest2x(h), pol.w(P[f"{pre}upsamples.{li}.1.weight"]),
                            P[f"{pre}upsamples.{li}.1.bias"], 1, 1)
            h = pol.act(h_full)
            tap(f"upsamples.{li}", h)

    a = pol.act(silu(group_norm(h, P[pre + "output_conv.0.weight"], P[pre + "output_conv.0.bias"],
                                stats_from=h_full)))
    # the 64->3 output conv runs on tensor cores (bf16 weights, fp32 accumulate) and writes fp32
    return conv2d(a, pol.w(P[pre + "output_conv.2.weight"]), P[pre + "output_conv.2.bias"], 1, 1)


# ----------------------------------------------------------------------------------------------------------
# flow-model level
# ----------------------------------------------------------------------------------------------------------
def euler_sample(P, noise: np.ndarray, num_steps: int, spec: Optional[UNetSpec] = None,
                 policy: Policy = FP32_POLICY, return_trajectory: bool = False, save_every: int = 1):
    """BaseFlowModel.sample / sample_with_trajectory, models/base_flow.py:158-177,196-208.

    dt = 1/num_steps is a Python double; t_i = i*dt is rounded to fp32 when the tensor is built
    (base_flow.py:164) and v*dt multiplies an fp32 tensor by a Python scalar (base_flow.py:170)."""
    x = np.asarray(noise, dtype=F32)
    dt = 1.0 / num_steps
    traj = [x.copy()]
    for i in range(num_steps):
        t = np.full((x.shape[0],), i * dt, dtype=F32)
        v = unet_forward(P, x, t, spec, policy)
        x = (x + v * F32(dt)).astype(F32)
        if (i + 1) % save_every == 0:
            traj.append(x.copy())
    return traj if return_trajectory else x


def get_interpolation(x0: np.ndarray, x1: np.ndarray, t: np.ndarray):
    """models/base_flow.py:81-89: x_t = (1-t) x0 + t x1, target = x1 - x0."""
    tt = np.asarray(t, dtype=F32).reshape(-1, 1, 1, 1)
    return ((1 - tt) * x0 + tt * x1).astype(F32), (x1 - x0).astype(F32)


def fm_loss(P, x0, x1, t, spec=None, policy: Policy = FP32_POLICY) -> float:
    """Loop body of train_rectified_flow up to the loss, models/rectified_flow.py:222-231 (eval-mode network:
    the oracle has no dropout, parity runs use dropout=0.0)."""
    xt, target = get_interpolation(x0, x1, t)
    pred = unet_forward(P, xt, t, spec, policy)
    return float(np.mean((pred.astype(np.float64) - target) ** 2))


def straightness(P, x0, x1, num_points: int = 10, spec=None, policy: Policy = FP32_POLICY) -> float:
    """RectifiedFlowModel.compute_straightness, models/rectified_flow.py:98-124."""
    ideal = (x1 - x0).astype(F32)
    x = np.array(x0, dtype=F32)
    dt = 1.0 / num_points
    dev = []
    for i in range(num_points):
        t = np.full((x.shape[0],), i * dt, dtype=F32)
        v = unet_forward(P, x, t, spec, policy)
        dev.append(float(np.mean((v.astype(np.float64) - ideal) ** 2)))
        x = (x + v * F32(dt)).astype(F32)
    return float(np.mean(dev))


def reflow_pairs(P, noise: np.ndarray, num_steps: int, batch_size: int = 32, spec=None,
                 policy: Policy = FP32_POLICY):
    """generate_reflow_pairs, models/rectified_flow.py:150-170, with the noise supplied by the caller
    (north star: host-seeded noise) instead of drawn per batch on the device."""
    outs = []
    for s in range(0, noise.shape[0], batch_size):
        outs.append(euler_sample(P, noise[s:s + batch_size], num_steps, spec, policy))
    return np.asarray(noise, dtype=F32), np.concatenate(outs, axis=0)


def unet_flops_per_image(spec: Optional[UNetSpec] = None, size: int = 64) -> float:
    """2*MACs of every conv / linear + 4*C*N^2 attention: 12.7636 GF @64², default."""
    spec = spec or UNetSpec()
    mc, chans, nres = spec.model_channels, spec.level_channels, spec.num_res_blocks
    td = spec.time_dim
    fl = 2.0 * (mc * td + td * td)
    conv = lambda ci, co, k, hw: 2.0 * ci * co * k * k * hw * hw
    res = lambda ci, co, hw: (conv(ci, co, 3, hw) + conv(co, co, 3, hw) + (conv(ci, co, 1, hw) if ci != co else 0)
                              + 2.0 * td * co)
    fl += conv(spec.in_channels, mc, 3, size)
    hw, cin = size, mc
    nlev = len(chans)
    for lv in range(nlev):
        for _ in range(nres):
            fl += res(cin, chans[lv], hw)
            cin = chans[lv]
        if lv < nlev - 1:
            fl += conv(cin, cin, 3, hw // 2)
            hw //= 2
    fl += 2 * res(cin, cin, hw)
    fl += conv(cin, 3 * cin, 1, hw) + conv(cin, cin, 1, hw) + 4.0 * cin * (hw * hw) ** 2
    for li, lv in enumerate(range(nlev - 1, -1, -1)):
        fl += res(cin + chans[lv], chans[lv], hw)
        for _ in range(nres - 1):
            fl += res(chans[lv], chans[lv], hw)
        cin = chans[lv]
        if lv > 0:
            hw *= 2
            fl += conv(cin, cin, 3, hw)
    fl += conv(chans[0], spec.out_channels, 3, hw)
    return fl
