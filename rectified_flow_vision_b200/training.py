"""Native training step: host-side driver of rfv_train_accumulate / rfv_optimizer_step.

Replaces the body of the reference training loops -- ``loss = mse(model(x_t, t), x1 - x0); optimizer.zero_grad();
loss.backward(); clip_grad_norm_(params, 1.0); optimizer.step()`` (models/rectified_flow.py:217-238,
models/base_flow.py:266-275) with torch.optim.AdamW(lr) defaults and a per-epoch CosineAnnealingLR
(models/rectified_flow.py:208-209) -- by one forward+backward through the CUDA engine and one fused
clip + AdamW kernel.  No autograd graph is built and nothing falls back to PyTorch.

Data parallel: every rank holds an identical replica, runs the step on its shard of the batch and the
flat gradient buffer is summed with ONE ``all_reduce`` (NCCL over NVLink on the GPU box, gloo in the CPU tests of the
sharding logic) before the identical optimizer step on every rank; ``grad_scale = 1 / world_size`` turns the sum of
per-rank means into the global mean (equal shards).
"""
from __future__ import annotations

import math
from typing import Optional

import torch


def cosine_lr(base_lr: float, epoch: int, epochs: int, eta_min: float = 0.0) -> float:
    """torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=epochs) closed form, stepped once per epoch."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / epochs)) / 2.0


class NativeTrainer:
    """AdamW state + step counter for one model (the optimizer object of the reference loops)."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, micro_batch: Optional[int] = None, process_group=None):
        self.model = model
        self.lr = float(lr)
        self.betas, self.eps, self.weight_decay, self.max_grad_norm = betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self._fresh = True      # a new trainer = a new torch.optim.AdamW: zero moments on first use of the engine
        self.process_group = process_group
        self._micro_batch = micro_batch
        self.last_grad_norm = None
        self._comm_stream = None

    def _engine(self, size: int):
        return self.model.velocity_net.train_engine(size, torch.device(self.model.device), self._micro_batch)

    def _rank(self) -> int:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.process_group)
        return 0

    def _world(self) -> int:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def _all_reduce_gradients(self, eng) -> None:
        """SUM over the ranks of the flat gradient buffer.  On NCCL the buffer is reduced bucket by bucket on a side stream: the
        engine lays the slots out in the order they become final and exposes one event pair per bucket, so the collective of
        a finished range runs under the rest of the backward pass (everything is enqueued asynchronously: the backward pass
        is already in the stream queues when the first wait is placed).  RFV_BUCKETS=0 or a non-NCCL group: one call."""
        import os
        import torch.distributed as dist
        buf = eng.grad_buffer()
        buckets = eng.grad_buckets() if hasattr(eng, "grad_buckets") else []
        overlap = (len(buckets) > 1 and buf.is_cuda and dist.get_backend(self.process_group) == "nccl"
                   and os.environ.get("RFV_BUCKETS", "1") != "0")
        if not overlap:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.process_group)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(buf.device)
        cur = torch.cuda.current_stream(buf.device)
        for k, (off, n) in enumerate(buckets):
            eng.grad_bucket_wait(k, self._comm_stream)
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(buf[off:off + n], op=dist.ReduceOp.SUM, group=self.process_group)
        cur.wait_stream(self._comm_stream)

    def step(self, x0: torch.Tensor, x1: torch.Tensor, t: torch.Tensor, lr: Optional[float] = None,
             dropout: Optional[float] = None, seed: Optional[int] = None) -> torch.Tensor:
        """One optimizer step on this rank's (x0, x1, t) shard; returns the shard's mean loss (0-dim device tensor)."""
        eng = self._engine(x0.shape[-1])
        p = self.model.velocity_net.dropout_p if dropout is None else float(dropout)
        if not self.model.training:
            p = 0.0  # nn.Dropout is the identity in eval mode
        if self._fresh:
            # the AdamW moments live in the engine cached on the UNet; every train_* call of the reference builds a new
            # optimizer with zero state (models/rectified_flow.py:208, models/base_flow.py:255)
            eng.reset_optimizer()
            self._fresh = False
        self.step_count += 1
        world = self._world()
        if seed is None:   # dropout stream: the step counter, with the rank in the high half (replicas draw different masks)
            seed = self.step_count | (self._rank() << 32)
        eng.zero_grad()
        loss = eng.train_accumulate(x0, x1, t, dropout_p=p, seed=seed)
        if world > 1:
            self._all_reduce_gradients(eng)
        self.last_grad_norm = eng.optimizer_step(self.lr if lr is None else lr, self.step_count, self.betas[0],
                                                 self.betas[1], self.eps, self.weight_decay, self.max_grad_norm,
                                                 grad_scale=1.0 / world)
        return loss
