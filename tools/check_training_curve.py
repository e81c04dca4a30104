"""Training-dynamics parity: N optimizer steps of the native trainer vs the same steps through PyTorch autograd +
torch.optim.AdamW on the functional port (same seeded weights, same seeded batches, dropout off), both on the GPU.
Prints the two loss curves and the relative L2 distance of the final parameters."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rectified_flow_vision_b200 as pkg
from rectified_flow_vision_b200.training import NativeTrainer
from oracle import torch_port

STEPS, B, LR = int(os.environ.get("STEPS", "60")), 32, 2e-4
dev = "cuda:0"
kw = dict(image_size=32, channel_mult=[1, 2], num_res_blocks=1)
arch = dict(model_channels=64, channel_mult=(1, 2), num_res_blocks=1)
torch.manual_seed(11)
m = pkg.RectifiedFlowModel(device=dev, **kw)
m.eval()                                       # dropout off in both arms
P = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
p0 = {k: v.detach().clone() for k, v in P.items()}
opt = torch.optim.AdamW(list(P.values()), lr=LR)
tr = NativeTrainer(m, lr=LR, micro_batch=B)
g = torch.Generator().manual_seed(3)
data_x1 = torch.tanh(torch.randn(256, 3, 32, 32, generator=g)) * 0.5 + torch.linspace(-0.5, 0.5, 32).view(1, 1, 1, 32)   # structured targets
la, lb = [], []
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for step in range(STEPS):
    idx = torch.randint(0, 256, (B,), generator=g)
    x1 = data_x1[idx].to(dev)
    x0 = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    t = torch.rand(B, generator=g).to(dev)
    la.append(float(tr.step(x0, x1, t).item()))
    tt = t.view(-1, 1, 1, 1)
    with torch.device(dev):
        pred = torch_port.unet_forward_grad(P, (1 - tt) * x0 + tt * x1, t, **arch)
    loss = torch.nn.functional.mse_loss(pred, x1 - x0)
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
    opt.step()
    lb.append(float(loss.item()))
la, lb = np.array(la), np.array(lb)
sd = dict(m.named_parameters())
num = sum(float(((sd[k].detach() - P[k].detach()) ** 2).sum()) for k in P)
den = sum(float(((P[k].detach() - p0[k]) ** 2).sum()) for k in P)
print("step   native    torch-fp32")
for s in list(range(0, STEPS, max(1, STEPS // 12))) + [STEPS - 1]:
    print(f"{s:4d}  {la[s]:8.4f}  {lb[s]:8.4f}")
print(f"mean |loss difference| / mean loss over {STEPS} steps: {np.abs(la - lb).mean() / lb.mean():.3e}")
print(f"final parameters: ||native - torch|| / ||torch - init|| = {(num / den) ** 0.5:.3e}")
print(f"loss went from {lb[:5].mean():.4f} to {lb[-5:].mean():.4f} (torch) / {la[:5].mean():.4f} to {la[-5:].mean():.4f} (native)")
