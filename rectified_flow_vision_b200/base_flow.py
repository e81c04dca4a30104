"""BaseFlowModel: host-side mirror of the reference ``models/base_flow.py`` over the sm_100a engine.

Same constructor, attributes, method names, argument meaning, return types and checkpoint format as the
reference (``models/base_flow.py:24-226``); the numerical bodies are single calls into the C-ABI library:

  forward                -> rfv_velocity        (replaces models/base_flow.py:102 -> models/unet.py:229-275)
  sample                 -> rfv_euler_sample    (replaces the Python Euler loop, models/base_flow.py:163-173)
  sample_with_trajectory -> rfv_euler_sample    (models/base_flow.py:196-208)
  compute_loss           -> rfv_fm_loss         (models/base_flow.py:113-129; forward value only, see below)
  save / load            -> unchanged torch.save / torch.load of {'state_dict','config'} (models/base_flow.py:210-226)

``train_base_flow`` / ``train_rectified_flow`` drive the native step (``training.NativeTrainer``: rfv_train_accumulate +
rfv_optimizer_step, no graph).  A caller-written loop over ``compute_loss(x).backward()`` / ``model(x_t, t)`` in training
mode with any ``torch.optim`` optimizer works too: ``autograd.py`` wraps rfv_train_forward / rfv_train_backward in a
``torch.autograd.Function`` that fills ``.grad`` of the 174 parameters.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .unet import UNet


class BaseFlowModel(nn.Module):
    def __init__(self, image_size: int = 64, in_channels: int = 3, model_channels: int = 64,
                 channel_mult: List[int] = [1, 2, 4], num_res_blocks: int = 2,
                 attention_resolutions: List[int] = [16, 8], dropout: float = 0.1,
                 device: str = 'cuda' if torch.cuda.is_available() else 'cpu'):
        super().__init__()
        self.image_size = image_size
        self.in_channels = in_channels
        self.device = device
        self.velocity_net = UNet(in_channels=in_channels, model_channels=model_channels, out_channels=in_channels,
                                 channel_mult=channel_mult, num_res_blocks=num_res_blocks,
                                 attention_resolutions=attention_resolutions, dropout=dropout)
        self.to(device)

    # ----- elementwise helpers (kept in torch: they are the caller-side glue of the reference API) --------
    def get_interpolation(self, x0: torch.Tensor, x1: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x_t = (1-t) x0 + t x1, target = x1 - x0 (models/base_flow.py:81-89)."""
        t = t.view(-1, 1, 1, 1)
        return (1 - t) * x0 + t * x1, x1 - x0

    # ----- engine access -----------------------------------------------------------------------------------
    def _engine(self, size: Optional[int] = None):
        return self.velocity_net.engine(size or self.image_size, torch.device(self.device))

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return self.velocity_net(x, t)

    def compute_loss(self, x1: torch.Tensor) -> torch.Tensor:
        """Flow-matching loss for a data batch (models/base_flow.py:104-131): fresh x0 ~ N(0,I), t ~ U[0,1).
        In training mode the result carries an autograd graph to every parameter (``loss.backward()`` runs the native
        backward pass and fills ``.grad``, see ``autograd.py``); in eval mode it is the fused forward-only kernel path."""
        x0 = torch.randn_like(x1)
        t = torch.rand(x1.shape[0], device=x1.device)
        if self.training:
            x_t, target = self.get_interpolation(x0, x1, t)
            return torch.nn.functional.mse_loss(self.forward(x_t, t), target)
        return self._engine(x1.shape[-1]).fm_loss(x0, x1, t)

    @torch.no_grad()
    def sample(self, noise: Optional[torch.Tensor] = None, num_steps: int = 100, batch_size: int = 1,
               return_trajectory: bool = False):
        """N-step Euler integration from noise (t=0) to data (t=1); models/base_flow.py:133-177."""
        self.eval()
        if noise is None:
            noise = torch.randn(batch_size, self.in_channels, self.image_size, self.image_size, device=self.device)
        eng = self._engine(noise.shape[-1])
        if return_trajectory:
            x, traj = eng.euler_sample(noise, num_steps, save_every=1)
            return [noise] + ([traj[i] for i in range(traj.shape[0])] if traj is not None else [])
        x, _ = eng.euler_sample(noise, num_steps)
        return x

    @torch.no_grad()
    def sample_with_trajectory(self, noise: torch.Tensor, num_steps: int = 100, save_every: int = 10) -> List[torch.Tensor]:
        """models/base_flow.py:179-208: snapshots after every ``save_every``-th step, preceded by the noise."""
        self.eval()
        x, traj = self._engine(noise.shape[-1]).euler_sample(noise, num_steps, save_every=save_every)
        out = [noise.clone()]
        if traj is not None:
            out += [traj[i] for i in range(traj.shape[0])]
        return out

    # ----- checkpoint I/O: format unchanged ------------------------------------------------------------------
    def save(self, path: str):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save({'state_dict': self.state_dict(),
                    'config': {'image_size': self.image_size, 'in_channels': self.in_channels}}, path)
        print(f"Modelo guardado en: {path}")

    def load(self, path: str):
        checkpoint = torch.load(path, map_location=self.device)
        self.load_state_dict(checkpoint['state_dict'])
        print(f"Modelo cargado desde: {path}")


def train_base_flow(model: BaseFlowModel, dataloader, epochs: int = 50, lr: float = 1e-4,
                    save_path: Optional[str] = None, save_every: int = 10) -> List[float]:
    """models/base_flow.py:229-295: flow matching on real data, x0 ~ N(0, I) and t ~ U[0,1) drawn per batch
    (``compute_loss``, :113-129); step body on the native engine (see ``training.NativeTrainer``)."""
    import numpy as np
    from .training import NativeTrainer, cosine_lr

    trainer = NativeTrainer(model, lr=lr)
    losses: List[float] = []
    for epoch in range(epochs):
        model.train()
        cur_lr = cosine_lr(lr, epoch, epochs)
        epoch_losses = []
        for batch in dataloader:
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            x = x.to(model.device)
            x0 = torch.randn_like(x)
            t = torch.rand(x.shape[0], device=x.device)
            epoch_losses.append(trainer.step(x0, x, t, lr=cur_lr))
        avg_loss = float(torch.stack(epoch_losses).mean().item())
        losses.append(avg_loss)
        print(f"Epoch {epoch+1}/{epochs} - Loss: {avg_loss:.4f}")
        if save_path and (epoch + 1) % save_every == 0:
            model.save(f"{save_path}_epoch{epoch+1}.pt")
    if save_path:
        model.save(f"{save_path}_final.pt")
    return losses
