"""RectifiedFlowModel + reflow pair generation over the sm_100a engine.

Mirrors ``models/rectified_flow.py`` of the reference: ``RectifiedFlowModel`` (``:29-124``),
``generate_reflow_pairs`` (``:127-174``), ``train_rectified_flow`` (``:177-255``), ``iterative_reflow``
(``:258-318``) -- same names, arguments, defaults and return types.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .base_flow import BaseFlowModel


class RectifiedFlowModel(BaseFlowModel):
    def __init__(self, image_size: int = 64, in_channels: int = 3, model_channels: int = 64,
                 channel_mult: List[int] = [1, 2, 4], num_res_blocks: int = 2,
                 attention_resolutions: List[int] = [16, 8], dropout: float = 0.1,
                 device: str = 'cuda' if torch.cuda.is_available() else 'cpu'):
        super().__init__(image_size=image_size, in_channels=in_channels, model_channels=model_channels,
                         channel_mult=channel_mult, num_res_blocks=num_res_blocks,
                         attention_resolutions=attention_resolutions, dropout=dropout, device=device)
        self.reflow_iteration = 0

    @staticmethod
    def from_base_model(base_model: BaseFlowModel) -> 'RectifiedFlowModel':
        """Like the reference (models/rectified_flow.py:65-80): copies only image_size / in_channels / device;
        the architecture falls back to the defaults and weights are NOT copied."""
        return RectifiedFlowModel(image_size=base_model.image_size, in_channels=base_model.in_channels,
                                  device=base_model.device)

    def compute_straightness(self, x0: torch.Tensor, x1: torch.Tensor, num_points: int = 10) -> float:
        """Mean over the Euler steps of mse(v(x_t, t), x1 - x0) (models/rectified_flow.py:82-124).  The per-step
        MSEs are accumulated on the device; one read-back at the end instead of one ``.item()`` per step."""
        self.eval()
        with torch.no_grad():
            dev = self._engine(x0.shape[-1]).straightness(x0, x1, num_points)
        return float(np.mean(dev.cpu().numpy().astype(np.float64)))


def generate_reflow_pairs(teacher_model: BaseFlowModel, num_pairs: int, batch_size: int = 32,
                          num_steps: int = 100, noise: Optional[torch.Tensor] = None,
                          seed: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(noise, teacher sample) pairs as two CPU fp32 tensors [num_pairs, C, S, S] (models/rectified_flow.py:127-174).

    The reference draws each batch's noise on the device, unseeded (``:159-161``).  Here the noise comes from
    the HOST: pass ``noise`` (CPU fp32 [num_pairs,C,S,S]) or a ``seed`` for a CPU generator; with neither, the
    global CPU generator is used.  ``batch_size`` only bounds how many images are integrated together on the
    device (the engine's micro-batch); results do not depend on it.  Under ``torch.distributed`` use
    ``rectified_flow_vision_b200.dist.generate_reflow_pairs_sharded`` to split the pairs across ranks."""
    teacher_model.eval()
    c, s = teacher_model.in_channels, teacher_model.image_size
    if noise is None:
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        noise = torch.randn(num_pairs, c, s, s, generator=gen)
    if noise.shape[0] != num_pairs:
        raise ValueError("noise.shape[0] must equal num_pairs")
    print(f"Generating {num_pairs} pairs for Reflow...")
    x0_all = noise.to(torch.float32).cpu().contiguous()
    if torch.cuda.is_available() and not x0_all.is_pinned():
        x0_all = x0_all.pin_memory()
    x1_all = teacher_model._engine(s).euler_sample_host(x0_all, num_steps)
    print(f"Generated {x0_all.shape[0]} pairs")
    return x0_all, x1_all


def train_rectified_flow(model: RectifiedFlowModel, x0_data: torch.Tensor, x1_data: torch.Tensor,
                         epochs: int = 30, batch_size: int = 16, lr: float = 1e-4,
                         save_path: Optional[str] = None, save_every: int = 10) -> List[float]:
    """models/rectified_flow.py:177-255 with the step body (interpolation, forward, MSE, backward, clip, AdamW) run by
    the native engine: same shuffled TensorDataset/DataLoader batching, t ~ U[0,1) drawn on the device per batch,
    torch.optim.AdamW(lr) defaults, CosineAnnealingLR(epochs) stepped per epoch, checkpoints every ``save_every``."""
    from torch.utils.data import DataLoader, TensorDataset
    from .training import NativeTrainer, cosine_lr

    dataset = TensorDataset(x0_data, x1_data)
    dataloader = DataLoader(dataset, batch_size=batch_size, shuffle=True)
    trainer = NativeTrainer(model, lr=lr)
    losses: List[float] = []
    for epoch in range(epochs):
        model.train()
        cur_lr = cosine_lr(lr, epoch, epochs)
        epoch_losses = []
        for x0, x1 in dataloader:
            x0 = x0.to(model.device)
            x1 = x1.to(model.device)
            t = torch.rand(x0.shape[0], device=model.device)
            epoch_losses.append(trainer.step(x0, x1, t, lr=cur_lr))
        avg_loss = float(torch.stack(epoch_losses).mean().item())  # one host sync per epoch, not per step
        losses.append(avg_loss)
        print(f"Reflow Epoch {epoch+1}/{epochs} - Loss: {avg_loss:.4f}")
        if save_path and (epoch + 1) % save_every == 0:
            model.save(f"{save_path}_epoch{epoch+1}.pt")
    if save_path:
        model.save(f"{save_path}_final.pt")
    return losses


def iterative_reflow(initial_model: BaseFlowModel, real_data_loader, num_iterations: int = 2,
                     epochs_per_iter: int = 30, num_pairs: int = 5000, teacher_steps: int = 100,
                     lr: float = 1e-4, save_dir: Optional[str] = None) -> List[RectifiedFlowModel]:
    """models/rectified_flow.py:258-318 (teacher -> pairs -> student, halving teacher steps each round)."""
    models: List[RectifiedFlowModel] = []
    teacher = initial_model
    for k in range(num_iterations):
        student = RectifiedFlowModel.from_base_model(teacher)
        student.reflow_iteration = k + 1
        x0, x1 = generate_reflow_pairs(teacher, num_pairs=num_pairs, num_steps=teacher_steps)
        train_rectified_flow(student, x0, x1, epochs=epochs_per_iter, lr=lr,
                             save_path=f"{save_dir}/reflow_k{k + 1}" if save_dir else None)
        models.append(student)
        teacher = student
        teacher_steps = max(teacher_steps // 2, 10)
    return models
