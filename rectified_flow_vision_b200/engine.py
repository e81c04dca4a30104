"""ctypes binding of include/rfv.h: the only door from Python into the CUDA path.

No torch types cross the boundary: tensors are passed as raw device pointers (``Tensor.data_ptr()``) plus
sizes, the stream as ``torch.cuda.current_stream().cuda_stream``.  If the shared library cannot be loaded the
import of this module still succeeds (so the parameter container / checkpoint I/O work on a CPU box) but any
attempt to compute raises -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _build

RFV_MAX_LEVELS = 8
FLAG_NO_UMMA = 1
FLAG_ONE_LANE = 2
FLAG_NO_GRAPH = 4194304
FLAG_KEEP_ACTS = 4
FLAG_TRAIN = 32
FLAG_NO_PDL = 16777216   # RFV_FLAG_NO_PDL


class RfvConfig(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("in_channels", C.c_int32), ("out_channels", C.c_int32),
                ("model_channels", C.c_int32), ("num_levels", C.c_int32),
                ("channel_mult", C.c_int32 * RFV_MAX_LEVELS), ("num_res_blocks", C.c_int32),
                ("num_heads", C.c_int32), ("micro_batch", C.c_int32), ("device", C.c_int32),
                ("flags", C.c_int32)]


class RfvAdamW(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("max_grad_norm", C.c_float), ("grad_scale", C.c_float),
                ("step", C.c_int64)]


class _DeviceArray:
    """Exposes engine-owned device memory to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


# every symbol include/rfv.h declares: (restype, argtypes)
_VP, _FP, _I64 = C.c_void_p, C.c_void_p, C.c_int64
SYMBOLS = {
    "rfv_abi_version": (C.c_int, []),
    "rfv_last_error": (C.c_char_p, []),
    "rfv_create": (C.c_int, [C.POINTER(RfvConfig), C.POINTER(_VP)]),
    "rfv_destroy": (C.c_int, [_VP]),
    "rfv_num_tensors": (C.c_int, [_VP]),
    "rfv_tensor_info": (C.c_int, [_VP, C.c_int, C.c_char_p, C.c_int, C.POINTER(_I64)]),
    "rfv_set_tensor": (C.c_int, [_VP, C.c_char_p, _FP, _I64, _VP]),
    "rfv_get_tensor": (C.c_int, [_VP, C.c_char_p, _FP, _I64, _VP]),
    "rfv_get_master": (C.c_int, [_VP, C.c_char_p, _FP, _I64, _VP]),
    "rfv_velocity": (C.c_int, [_VP, _FP, _FP, _FP, _I64, _VP]),
    "rfv_euler_sample": (C.c_int, [_VP, _FP, _I64, C.c_int, _FP, C.c_int, _VP]),
    "rfv_euler_sample_host": (C.c_int, [_VP, _FP, _FP, _I64, C.c_int]),
    "rfv_straightness": (C.c_int, [_VP, _FP, _FP, _I64, C.c_int, _FP, _VP]),
    "rfv_fm_loss": (C.c_int, [_VP, _FP, _FP, _FP, _I64, _FP, _VP]),
    "rfv_zero_grad": (C.c_int, [_VP, _VP]),
    "rfv_train_accumulate": (C.c_int, [_VP, _FP, _FP, _FP, _I64, C.c_float, C.c_uint64, _FP, _VP]),
    "rfv_train_forward": (C.c_int, [_VP, _FP, _FP, _I64, C.c_float, C.c_uint64, _FP, _VP]),
    "rfv_train_backward": (C.c_int, [_VP, _FP, _I64, _VP]),
    "rfv_reset_optimizer": (C.c_int, [_VP, _VP]),
    "rfv_grad_buffer": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_I64)]),
    "rfv_grad_bucket_count": (C.c_int, [_VP]),
    "rfv_grad_bucket_info": (C.c_int, [_VP, C.c_int, C.POINTER(_I64), C.POINTER(_I64)]),
    "rfv_grad_bucket_wait": (C.c_int, [_VP, C.c_int, _VP]),
    "rfv_get_grad": (C.c_int, [_VP, C.c_char_p, _FP, _I64, C.c_float, _VP]),
    "rfv_bind_param": (C.c_int, [_VP, C.c_char_p, _FP]),
    "rfv_optimizer_step": (C.c_int, [_VP, C.POINTER(RfvAdamW), _FP, _VP]),
    "rfv_metrics_mean": (C.c_int, [_FP, _I64, _I64, _VP, _VP]),
    "rfv_metrics_covariance": (C.c_int, [_FP, _VP, _I64, _I64, _VP, _VP]),
    "rfv_metrics_fid_terms": (C.c_int, [_FP, _VP, _I64, _FP, _VP, _I64, _I64, _VP, _VP, _VP]),
    "rfv_metrics_ssim": (C.c_int, [_FP, _FP, _I64, C.c_int, C.c_int, C.c_int, C.c_float, _VP, _VP]),
    "rfv_launch_count": (_I64, [_VP, C.c_int]),
    "rfv_flops_per_image": (C.c_double, [_VP]),
    "rfv_debug_activation": (_I64, [_VP, C.c_char_p, _FP, _I64, _VP]),
    "rfv_set_profiling": (C.c_int, [_VP, C.c_int]),
    "rfv_profile_report": (C.c_int, [_VP, C.c_char_p, C.c_int]),
}

_lib = None


def library_path() -> str:
    # RFV_LIB: an alternative build of the same sources (A/B measurements of compile-time switches on one GPU box)
    return os.environ.get("RFV_LIB") or str(_build.LIB)


def load_library():
    """dlopen librfv_b200.so and type every entry point.  Raises if it is missing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"native library {path} is missing: run `python __graft_entry__.py build` (needs nvcc). "
            "rectified_flow_vision_b200 has no non-CUDA fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.rfv_abi_version() != 1:
        raise RuntimeError("librfv_b200.so ABI version mismatch")
    _lib = lib
    return lib


class RfvError(RuntimeError):
    pass


def _check(rc: int):
    if rc != 0:
        msg = load_library().rfv_last_error()
        raise RfvError(f"rfv error {rc}: {msg.decode() if msg else '?'}")


def default_micro_batch(image_size: int) -> int:
    env = os.environ.get("RFV_MICRO_BATCH")
    if env:
        return int(env)
    return max(8, (512 * 64 * 64) // (image_size * image_size))  # ~2 GB of activations; +6 % throughput over 256 (measured)


class Engine:
    """One native handle: architecture + resolution + device.  Owns packed weights and the activation arena."""

    def __init__(self, arch: Dict, image_size: int, device: torch.device, micro_batch: Optional[int] = None,
                 flags: Optional[int] = None, train: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: rectified_flow_vision_b200 runs only on sm_100a GPUs")
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.image_size = int(image_size)
        self.in_channels = arch["in_channels"]
        cfg = RfvConfig()
        cfg.image_size = image_size
        cfg.in_channels = arch["in_channels"]
        cfg.out_channels = arch["out_channels"]
        cfg.model_channels = arch["model_channels"]
        mult = list(arch["channel_mult"])
        if len(mult) > RFV_MAX_LEVELS:
            raise ValueError("too many levels")
        cfg.num_levels = len(mult)
        for i, m in enumerate(mult):
            cfg.channel_mult[i] = m
        cfg.num_res_blocks = arch["num_res_blocks"]
        cfg.num_heads = 4
        cfg.micro_batch = micro_batch or default_micro_batch(image_size)
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.flags = int(os.environ.get("RFV_FLAGS", "0")) if flags is None else flags
        if os.environ.get("RFV_LANES", "2") == "1":
            cfg.flags |= FLAG_ONE_LANE      # A/B: one chain of micro-batches instead of two alternately enqueued ones
        if train:
            cfg.flags |= FLAG_TRAIN
        # Programmatic dependent launch pays on launch-bound micro-batches only; the library's bound is 128 images per call,
        # calibrated at 64x64.  A kernel over 128 images of 128x128 is as long as one over 512 of 64x64, where dependent launch
        # measured -1.2 %: scale the bound with the pixel count for larger images.
        if image_size > 64 and cfg.micro_batch * image_size * image_size > 128 * 64 * 64:
            cfg.flags |= FLAG_NO_PDL
        self.train = bool(cfg.flags & FLAG_TRAIN)
        self.micro_batch = cfg.micro_batch
        h = _VP()
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._cfg = cfg
        self._versions: Dict[str, tuple] = {}
        self.tensor_names = []
        buf = C.create_string_buffer(256)
        n = _I64()
        for i in range(self.lib.rfv_num_tensors(self.h)):
            _check(self.lib.rfv_tensor_info(self.h, i, buf, 256, C.byref(n)))
            self.tensor_names.append((buf.value.decode(), n.value))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.rfv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ----- helpers ---------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_f32(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if t.device.type != "cuda":
            raise ValueError(f"{name} must be a CUDA tensor on {self.device} (got {t.device}); there is no CPU path")
        if t.device != self.device:
            t = t.to(self.device)
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def _images(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """Validated fp32 device copy of an image batch: the C ABI only receives a row count and reads / writes
        B * in_channels * S * S floats, so a wrong channel count or a non-square tensor must be stopped here (the
        reference raises a shape error in its first conv)."""
        s = self.image_size
        if x.dim() != 4 or x.shape[0] < 1 or x.shape[1] != self.in_channels or x.shape[2] != s or x.shape[3] != s:
            raise ValueError(f"{name}: expected shape [B,{self.in_channels},{s},{s}] for this engine, got {tuple(x.shape)}")
        return self._dev_f32(x, name)

    def _times(self, t: torch.Tensor, batch: int) -> torch.Tensor:
        if t.dim() != 1 or t.shape[0] != batch:
            raise ValueError(f"t: expected shape [{batch}], got {tuple(t.shape)}")
        return self._dev_f32(t, "t")

    # ----- weights ---------------------------------------------------------------------------------------
    def sync_weights(self, unet: torch.nn.Module, prefix: str = "velocity_net.") -> None:
        """(Re)upload parameters whose storage or version counter changed since the last upload."""
        params = dict(unet.named_parameters())
        with torch.cuda.device(self.device):
            for full, numel in self.tensor_names:
                key = full[len(prefix):] if full.startswith(prefix) else full
                p = params[key]
                # _weights_gen: bumped when a native optimizer step rewrote the storage behind torch's back
                tag = (p.data_ptr(), p._version, getattr(unet, "_weights_gen", 0))
                if self._versions.get(full) == tag:
                    continue
                src = p.detach()
                if src.device != self.device or src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.to(self.device, torch.float32).contiguous()
                if src.numel() != numel:
                    raise ValueError(f"{full}: expected {numel} elements, got {src.numel()}")
                _check(self.lib.rfv_set_tensor(self.h, full.encode(), src.data_ptr(), numel, self._stream()))
                self._versions[full] = tag
                del src

    def get_tensor(self, name: str, numel: int) -> torch.Tensor:
        out = torch.empty(numel, dtype=torch.float32, device=self.device)
        _check(self.lib.rfv_get_tensor(self.h, name.encode(), out.data_ptr(), numel, self._stream()))
        return out

    # ----- hot path --------------------------------------------------------------------------------------
    def velocity(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        x = self._images(x, "x")
        t = self._times(t, x.shape[0])
        v = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_velocity(self.h, x.data_ptr(), t.data_ptr(), v.data_ptr(), x.shape[0],
                                         self._stream()))
        return v

    def euler_sample(self, noise: torch.Tensor, num_steps: int, save_every: int = 0):
        """Returns (x_final, traj or None); noise is not modified."""
        if int(num_steps) < 1:
            raise ValueError("num_steps must be >= 1")
        x = self._images(noise, "noise").clone()
        traj = None
        tp = None
        if save_every and save_every > 0 and num_steps // save_every > 0:
            traj = torch.empty((num_steps // save_every,) + tuple(x.shape), dtype=torch.float32, device=x.device)
            tp = traj.data_ptr()
        # batches beyond one micro-batch run as two alternately enqueued chains on two streams INSIDE the library
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_euler_sample(self.h, x.data_ptr(), x.shape[0], int(num_steps), tp,
                                             int(save_every or 0), self._stream()))
        return x, traj

    def euler_sample_host(self, noise_host: torch.Tensor, num_steps: int, out: Optional[torch.Tensor] = None):
        """Host fp32 [N,C,S,S] -> host fp32, H2D/D2H inside (pinned buffers are copied asynchronously)."""
        if noise_host.device.type != "cpu" or noise_host.dtype != torch.float32:
            raise ValueError("noise_host must be a CPU fp32 tensor")
        s_ = self.image_size
        if noise_host.dim() != 4 or noise_host.shape[0] < 1 or tuple(noise_host.shape[1:]) != (self.in_channels, s_, s_):
            raise ValueError(f"noise_host: expected shape [N,{self.in_channels},{s_},{s_}], got {tuple(noise_host.shape)}")
        if int(num_steps) < 1:
            raise ValueError("num_steps must be >= 1")
        noise_host = noise_host.contiguous()
        if out is not None and (out.device.type != "cpu" or out.dtype != torch.float32 or out.shape != noise_host.shape
                                or not out.is_contiguous()):
            raise ValueError("out must be a contiguous CPU fp32 tensor of the noise's shape")
        if out is None:
            out = torch.empty_like(noise_host, pin_memory=noise_host.is_pinned())
        torch.cuda.current_stream(self.device).synchronize()   # weight uploads were enqueued on the caller's stream
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_euler_sample_host(self.h, noise_host.data_ptr(), out.data_ptr(), noise_host.shape[0],
                                                  int(num_steps)))
        return out

    def straightness(self, x0: torch.Tensor, x1: torch.Tensor, num_points: int) -> torch.Tensor:
        x0 = self._images(x0, "x0")
        x1 = self._images(x1, "x1")
        if x1.shape != x0.shape:
            raise ValueError(f"shape mismatch: x0 {tuple(x0.shape)}, x1 {tuple(x1.shape)}")
        if int(num_points) < 1:
            raise ValueError("num_points must be >= 1")
        out = torch.empty(num_points, dtype=torch.float32, device=x0.device)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_straightness(self.h, x0.data_ptr(), x1.data_ptr(), x0.shape[0], int(num_points),
                                             out.data_ptr(), self._stream()))
        return out

    def fm_loss(self, x0: torch.Tensor, x1: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        x0 = self._images(x0, "x0")
        x1 = self._images(x1, "x1")
        if x1.shape != x0.shape:
            raise ValueError(f"shape mismatch: x0 {tuple(x0.shape)}, x1 {tuple(x1.shape)}")
        t = self._times(t, x0.shape[0])
        out = torch.empty((), dtype=torch.float32, device=x0.device)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_fm_loss(self.h, x0.data_ptr(), x1.data_ptr(), t.data_ptr(), x0.shape[0],
                                        out.data_ptr(), self._stream()))
        return out

    # ----- training (needs train=True) --------------------------------------------------------------------
    def zero_grad(self) -> None:
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_zero_grad(self.h, self._stream()))

    def train_accumulate(self, x0: torch.Tensor, x1: torch.Tensor, t: torch.Tensor, dropout_p: float = 0.0,
                         seed: int = 0) -> torch.Tensor:
        """loss = mean((v((1-t) x0 + t x1, t) - (x1 - x0))^2); parameter gradients are ADDED to the flat
        gradient buffer.  Returns the loss as a 0-dim device tensor (no host sync)."""
        x0 = self._images(x0, "x0")
        x1 = self._images(x1, "x1")
        if x0.shape != x1.shape:
            raise ValueError(f"shape mismatch: x0 {tuple(x0.shape)}, x1 {tuple(x1.shape)}")
        t = self._times(t, x0.shape[0])
        out = torch.empty((), dtype=torch.float32, device=x0.device)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_train_accumulate(self.h, x0.data_ptr(), x1.data_ptr(), t.data_ptr(), x0.shape[0],
                                                 float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(),
                                                 self._stream()))
        return out

    def train_forward(self, x: torch.Tensor, t: torch.Tensor, dropout_p: float, seed: int, out: torch.Tensor) -> None:
        """v = velocity_net(x, t) in training mode for ONE micro-batch, activations kept for ``train_backward``.
        x, t, out must already be validated fp32 device tensors (``_images`` / ``_times``); x and t must outlive the
        backward call."""
        if x.shape[0] > self.micro_batch:
            raise ValueError(f"train_forward keeps one micro-batch: batch {x.shape[0]} > micro_batch {self.micro_batch}")
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_train_forward(self.h, x.data_ptr(), t.data_ptr(), x.shape[0], float(dropout_p),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(), self._stream()))
        self.fwd_token = getattr(self, "fwd_token", 0) + 1

    def train_backward(self, dv: torch.Tensor) -> None:
        """Adds dL/dparam for the micro-batch of the last ``train_forward`` into the flat gradient buffer."""
        dv = self._images(dv, "dv")
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_train_backward(self.h, dv.data_ptr(), dv.shape[0], self._stream()))
        self.fwd_token = getattr(self, "fwd_token", 0) + 1   # the kept activations are consumed

    def reset_optimizer(self) -> None:
        """Zero the AdamW moments (what constructing a fresh torch.optim.AdamW does)."""
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_reset_optimizer(self.h, self._stream()))

    def grad_buffer(self) -> torch.Tensor:
        """The engine's flat fp32 gradient buffer as a torch tensor sharing its memory (all-reduce target)."""
        ptr, n = _VP(), _I64()
        _check(self.lib.rfv_grad_buffer(self.h, C.byref(ptr), C.byref(n)))
        return torch.as_tensor(_DeviceArray(ptr.value, n.value), device=self.device)

    def grad_buckets(self):
        """[(offset, numel)] of the contiguous ranges of the flat gradient buffer, in the order they become final."""
        out = []
        off, n = _I64(), _I64()
        for k in range(int(self.lib.rfv_grad_bucket_count(self.h))):
            _check(self.lib.rfv_grad_bucket_info(self.h, k, C.byref(off), C.byref(n)))
            out.append((off.value, n.value))
        return out

    def grad_bucket_wait(self, index: int, stream: "torch.cuda.Stream") -> None:
        """`stream` waits until bucket `index` of the last ``train_accumulate`` is final."""
        _check(self.lib.rfv_grad_bucket_wait(self.h, int(index), C.c_void_p(stream.cuda_stream)))

    def get_grad(self, name: str, numel: int, scale: float = 1.0) -> torch.Tensor:
        out = torch.empty(numel, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_get_grad(self.h, name.encode(), out.data_ptr(), numel, float(scale), self._stream()))
        return out

    def bind_params(self, unet: torch.nn.Module, prefix: str = "velocity_net.") -> None:
        """The optimizer step then also writes the updated fp32 values into the module's own parameter storage, so
        ``state_dict()`` / ``save()`` always see the trained weights."""
        self.sync_weights(unet, prefix)
        params = dict(unet.named_parameters())
        self._bound = []
        self._bound_module, self._bound_prefix = unet, prefix
        for full, numel in self.tensor_names:
            key = full[len(prefix):] if full.startswith(prefix) else full
            p = params[key]
            if p.device != self.device or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"{full}: parameters must be contiguous fp32 tensors on {self.device} to be trained")
            _check(self.lib.rfv_bind_param(self.h, full.encode(), p.data_ptr()))
            self._bound.append(p)
        self._bound_ptrs = [p.data_ptr() for p in self._bound]

    def bound_storage_moved(self) -> bool:
        """True when a bound Parameter's storage was reallocated since ``bind_params`` (module.to(), p.data = ...): the
        optimizer would keep writing to the old address."""
        params = dict(self._bound_module.named_parameters())
        cur = []
        for full, _ in self.tensor_names:
            key = full[len(self._bound_prefix):] if full.startswith(self._bound_prefix) else full
            cur.append(params[key].data_ptr())
        return cur != getattr(self, "_bound_ptrs", None)

    def optimizer_step(self, lr: float, step: int, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                       weight_decay: float = 0.01, max_grad_norm: float = 1.0, grad_scale: float = 1.0) -> torch.Tensor:
        """clip_grad_norm_(max_grad_norm) + torch.optim.AdamW step; returns the pre-clip gradient norm (device)."""
        hp = RfvAdamW(lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale, int(step))
        out = torch.empty((), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib.rfv_optimizer_step(self.h, C.byref(hp), out.data_ptr(), self._stream()))
        # the bound parameter storage was rewritten behind torch's back: other engines of the module must re-upload,
        # this one already holds the new values
        unet = getattr(self, "_bound_module", None)
        if unet is not None:
            unet._weights_gen = getattr(unet, "_weights_gen", 0) + 1
            params = dict(unet.named_parameters())
            for full, _ in self.tensor_names:
                key = full[len(self._bound_prefix):] if full.startswith(self._bound_prefix) else full
                p = params[key]
                self._versions[full] = (p.data_ptr(), p._version, unet._weights_gen)
        return out

    # ----- introspection ---------------------------------------------------------------------------------
    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.rfv_launch_count(self.h, 1 if reset else 0))   # both lanes

    def flops_per_image(self) -> float:
        return float(self.lib.rfv_flops_per_image(self.h))

    def debug_activation(self, name: str, capacity: int) -> torch.Tensor:
        out = torch.empty(capacity, dtype=torch.float32, device=self.device)
        n = self.lib.rfv_debug_activation(self.h, name.encode(), out.data_ptr(), capacity, self._stream())
        if n < 0:
            _check(int(n))
        return out[:n]

    def set_profiling(self, on: bool):
        _check(self.lib.rfv_set_profiling(self.h, 1 if on else 0))

    def profile_report(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        _check(self.lib.rfv_profile_report(self.h, buf, len(buf)))
        return buf.value.decode()
