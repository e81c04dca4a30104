"""BASELINE.json configs[4]: config.yaml UNet at 128x128, random-init seeded weights, 8-step Euler, batch 128 per GPU.
Checks one velocity evaluation against the CPU port (fp32) on 2 images and times the 8-step sampler."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rectified_flow_vision_b200 as pkg
from oracle import torch_port
from tests import util

torch.manual_seed(0)
m = pkg.BaseFlowModel(image_size=128, device="cuda:0")
m.eval()
g = torch.Generator().manual_seed(42)
x = torch.randn(2, 3, 128, 128, generator=g)
t = torch.rand(2, generator=g)
P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
ref = torch_port.unet_forward(P, x, t)
v = m(x.cuda(), t.cuda()).cpu()
print("128x128 velocity rel-L2 vs fp32 port:", util.rel_l2(v.numpy(), ref.numpy()), "max-rel", util.max_rel(v.numpy(), ref.numpy()))
B = 128
nz = torch.randn(B, 3, 128, 128, generator=g).cuda()
eng = m._engine(128)
for _ in range(2):
    eng.euler_sample(nz, 8)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    eng.euler_sample(nz, 8)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"8-step Euler, batch {B} @128x128: {ms:.2f} ms -> {B / ms * 1e3:.1f} images/s; {B * 8 * 51.8555e9 / ms / 1e9:.1f} TFLOP/s; micro_batch {eng.micro_batch}")
# training at 128x128 (attention over 1024 tokens in the backward pass too)
from oracle import train_oracle as T
x1 = torch.randn(2, 3, 128, 128, generator=g)
loss_ref, grads = T.loss_and_grads(P, x, x1, t)
te = m.velocity_net.train_engine(128, "cuda:0", micro_batch=2)
te.zero_grad()
loss = float(te.train_accumulate(x.cuda(), x1.cuda(), t.cuda(), 0.0, 1).item())
worst = max((util.rel_l2(te.get_grad(k, gr.numel()).cpu().numpy().reshape(gr.shape), gr.numpy()), k) for k, gr in grads.items())
print(f"128x128 train step: loss {loss:.5f} (ref {loss_ref:.5f}); worst gradient rel-L2 {worst[0]:.3e} at {worst[1]}")
