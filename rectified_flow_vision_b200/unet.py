"""UNet velocity field: parameter container + dispatch to the sm_100a engine.

Drop-in for the reference ``models/unet.py`` UNet (ctor ``models/unet.py:136-145``, forward ``:229-275``):
same constructor arguments, same 174 ``state_dict`` keys / shapes / dtypes (fp32, OIHW), and -- because the
parameters are created with the same initialisers in the same order as the reference constructor --
``torch.manual_seed(s); UNet()`` yields bit-identical initial weights to the reference under the same seed
(checked by ``oracle/make_golden.py`` and ``tests/test_api_cpu.py``).

Unlike the reference there is no layer code here: the module tree only holds ``nn.Parameter``s.  ``forward``
hands raw device pointers to the C-ABI library (``include/rfv.h``), which runs the hand-written CUDA path.
There is no CPU / PyTorch fallback: without a GPU or without the built library ``forward`` raises.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import engine as _engine


def _topology(in_ch: int, mc: int, out_ch: int, mult: Sequence[int], nres: int):
    """Walk the architecture once in the reference's CONSTRUCTION order (models/unet.py:157-227) and yield
    (key, kind, shape) for every parameterised layer.  kind: 'linear' | 'conv' | 'norm'."""
    td = 4 * mc
    chans = [mc * m for m in mult]
    nlev = len(chans)

    def res(prefix, ci, co):
        yield prefix + "norm1", "norm", (ci,)
        yield prefix + "conv1", "conv", (co, ci, 3, 3)
        yield prefix + "norm2", "norm", (co,)
        yield prefix + "conv2", "conv", (co, co, 3, 3)
        yield prefix + "time_mlp.1", "linear", (co, td)
        if ci != co:
            yield prefix + "shortcut", "conv", (co, ci, 1, 1)

    yield "time_mlp.1", "linear", (td, mc)
    yield "time_mlp.3", "linear", (td, td)
    yield "input_conv", "conv", (mc, in_ch, 3, 3)
    ci, bi = mc, 0
    for lv in range(nlev):
        for _ in range(nres):
            yield from res(f"enc_blocks.{bi}.", ci, chans[lv])
            ci, bi = chans[lv], bi + 1
        if lv < nlev - 1:
            yield f"downsamples.{lv}", "conv", (ci, ci, 3, 3)
    yield from res("mid_block1.", ci, ci)
    yield "mid_attn.norm", "norm", (ci,)
    yield "mid_attn.qkv", "conv", (3 * ci, ci, 1, 1)
    yield "mid_attn.proj", "conv", (ci, ci, 1, 1)
    yield from res("mid_block2.", ci, ci)
    bi = 0
    for li, lv in enumerate(range(nlev - 1, -1, -1)):
        yield from res(f"dec_blocks.{bi}.", ci + chans[lv], chans[lv])
        bi += 1
        for _ in range(nres - 1):
            yield from res(f"dec_blocks.{bi}.", chans[lv], chans[lv])
            bi += 1
        ci = chans[lv]
        if lv > 0:
            yield f"upsamples.{li}.1", "conv", (ci, ci, 3, 3)
    yield "output_conv.0", "norm", (chans[0],)
    yield "output_conv.2", "conv", (out_ch, chans[0], 3, 3)


# registration order of the reference's sub-modules: fixes the ORDER of state_dict keys (not needed for
# load_state_dict, kept so that saved checkpoints list keys like the reference's).
_TOP_ORDER = ["time_mlp", "input_conv", "enc_blocks", "downsamples", "mid_block1", "mid_attn", "mid_block2",
              "dec_blocks", "upsamples", "output_conv"]
_BLOCK_ORDER = ["norm1", "conv1", "norm2", "conv2", "time_mlp", "shortcut", "norm", "qkv", "proj"]


class _Node(nn.Module):
    """Bare container; exists only so that parameter names nest like the reference's."""

    def child(self, name: str) -> "_Node":
        if name not in self._modules:
            self.add_module(name, _Node())
        return self._modules[name]


class UNet(_Node):
    def __init__(self, in_channels: int = 3, model_channels: int = 64, out_channels: int = 3,
                 channel_mult: List[int] = [1, 2, 4], num_res_blocks: int = 2,
                 attention_resolutions: List[int] = [16, 8], dropout: float = 0.1):
        super().__init__()
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.channel_mult = list(channel_mult)
        self.num_levels = len(channel_mult)
        self.num_res_blocks = num_res_blocks
        self.dropout_p = float(dropout)
        # attention_resolutions is accepted and ignored, as in the reference (models/unet.py:143).
        self.channels = [model_channels * m for m in channel_mult]

        layers = list(_topology(in_channels, model_channels, out_channels, channel_mult, num_res_blocks))

        def rank(key: str) -> Tuple:
            parts = key.split(".")
            r = [_TOP_ORDER.index(parts[0])]
            for p in parts[1:]:
                r.append(int(p) if p.isdigit() else 100 + (_BLOCK_ORDER.index(p) if p in _BLOCK_ORDER else 50))
            return tuple(r)

        for key, _, _ in sorted(layers, key=lambda l: rank(l[0])):  # containers, registration order
            node = self
            for p in key.split("."):
                node = node.child(p)
        for key, kind, shape in layers:  # parameters, construction (RNG) order
            node = self
            for p in key.split("."):
                node = node.child(p)
            if kind == "norm":
                node.weight = nn.Parameter(torch.ones(shape))
                node.bias = nn.Parameter(torch.zeros(shape))
            else:  # nn.Linear / nn.Conv2d default init: kaiming_uniform(a=sqrt 5) then U(+-1/sqrt(fan_in))
                w = torch.empty(shape)
                nn.init.kaiming_uniform_(w, a=math.sqrt(5))
                fan_in = w[0].numel()
                b = torch.empty(shape[0])
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                nn.init.uniform_(b, -bound, bound)
                node.weight = nn.Parameter(w)
                node.bias = nn.Parameter(b)
        self._engine = None
        self._train_engine = None

    # ----- engine plumbing ------------------------------------------------------------------------------
    def arch(self) -> Dict:
        return dict(in_channels=self.in_channels, model_channels=self.model_channels,
                    out_channels=self.out_channels, channel_mult=self.channel_mult,
                    num_res_blocks=self.num_res_blocks)

    def engine(self, image_size: int, device) -> "_engine.Engine":
        """The per-(device, resolution) native handle; weights are (re)packed when parameters changed."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(
                "rectified_flow_vision_b200 has no CPU path: the velocity field runs only through the sm_100a "
                f"CUDA library (got device {dev}).")
        eng = self._engine
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if eng is None or eng.image_size != image_size or eng.device != dev:
            eng = _engine.Engine(self.arch(), image_size, dev)
            self._engine = eng
        eng.sync_weights(self)
        return eng

    def train_engine(self, image_size: int, device, micro_batch=None) -> "_engine.Engine":
        """Handle with the backward plan (RFV_FLAG_TRAIN); the optimizer writes through to this module's parameters."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"rectified_flow_vision_b200 trains only on sm_100a GPUs (got device {dev}); no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        eng = self._train_engine
        if eng is None or eng.image_size != image_size or eng.device != dev or \
                (micro_batch is not None and micro_batch != eng.micro_batch):
            mb = micro_batch or max(8, (128 * 64 * 64) // (image_size * image_size))
            self._train_engine = None   # release the old arena before the new one is allocated
            del eng
            eng = _engine.Engine(self.arch(), image_size, dev, micro_batch=mb, train=True)
            eng.bind_params(self)
            self._train_engine = eng
        elif eng.bound_storage_moved():
            eng.bind_params(self)       # parameter storage was reallocated: re-upload and re-bind the write-through pointers
        else:
            eng.sync_weights(self)
        return eng

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        if self.training:
            # nn.Dropout active + an autograd edge to every parameter (models/unet.py:62, base_flow.py:268-270): the native
            # training forward / backward behind a torch.autograd.Function
            from .autograd import velocity_with_grad
            return velocity_with_grad(self, x, t)
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected x of shape [B,{self.in_channels},S,S], got {tuple(x.shape)}")
        if t.dim() != 1 or t.shape[0] != x.shape[0]:
            raise ValueError(f"expected t of shape [{x.shape[0]}], got {tuple(t.shape)}")
        return self.engine(x.shape[-1], x.device).velocity(x, t)


def count_parameters(model: nn.Module) -> int:
    """models/unet.py:278-280."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
