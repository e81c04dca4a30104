"""Diagnosis: per-tensor gradient error of the native backward against the CPU training oracle on (a) seeded, (b) perturbed and
(c) natively trained weights of the default 64x64 net.  Prints the worst tensors with their share of the total gradient norm."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rectified_flow_vision_b200 as pkg  # noqa: E402
from oracle import train_oracle as T  # noqa: E402
from oracle.perturb import perturb_state_dict  # noqa: E402
from tests import util  # noqa: E402


def report(tag, m, xt, x1t, t, arch, size, repeat=2):
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    loss_ref, grads = T.loss_and_grads(sd, xt, x1t, t, **arch)
    gmax = max(float(g.norm()) for g in grads.values())
    eng = m.velocity_net.train_engine(size, "cuda:0")
    runs = []
    for r in range(repeat):
        eng.zero_grad()
        loss = float(eng.train_accumulate(xt.cuda(), x1t.cuda(), t.cuda(), dropout_p=0.0, seed=1).item())
        runs.append({k: eng.get_grad(k, g.numel()).cpu().numpy().reshape(g.shape) for k, g in grads.items()})
    rows = []
    for k, g in grads.items():
        e = util.rel_l2(runs[0][k], g.numpy())
        rr = util.rel_l2(runs[0][k], runs[1][k]) if repeat > 1 else 0.0
        rows.append((e, rr, float(g.norm()) / gmax, k))
    rows.sort(reverse=True)
    print(f"== {tag}: loss {loss:.5f} (oracle {loss_ref:.5f}), gmax {gmax:.4e}")
    for e, rr, share, k in rows[:12]:
        print(f"   rel-L2 {e:.3e}  run-to-run {rr:.3e}  |g|/gmax {share:.2e}  {k}")


def main():
    kw, size = dict(image_size=64), 64
    arch = util.arch_of(kw)
    gen = torch.Generator().manual_seed(9)
    xt, x1t, t = torch.randn(2, 3, size, size, generator=gen), torch.randn(2, 3, size, size, generator=gen), torch.rand(2, generator=gen)
    torch.manual_seed(5)
    m = pkg.RectifiedFlowModel(device="cuda:0", **kw)
    report("seeded", m, xt, x1t, t, arch, size)
    m2 = pkg.RectifiedFlowModel(device="cuda:0", **kw)
    m2.load_state_dict(perturb_state_dict(m.state_dict()))
    report("perturbed", m2, xt, x1t, t, arch, size)
    n = 16 * 12
    x0 = torch.randn(n, 3, size, size, generator=gen)
    x1 = (0.5 * torch.randn(n, 3, size, size, generator=gen)).clamp(-1, 1)
    pkg.train_rectified_flow(m, x0, x1, epochs=2, batch_size=16, lr=1e-3)
    m.eval()
    report("trained 24 steps", m, xt, x1t, t, arch, size)
    # in-distribution inputs for the trained model
    report("trained, in-distribution batch", m, x0[:2], x1[:2], t, arch, size)
    # the exact sequence of tests/test_gpu_weights.py::test_natively_trained_weights[64]
    torch.manual_seed(5)
    m = pkg.RectifiedFlowModel(device="cuda:0", **kw)
    gen = torch.Generator().manual_seed(9)
    x0 = torch.randn(n, 3, size, size, generator=gen)
    x1 = (0.5 * torch.randn(n, 3, size, size, generator=gen)).clamp(-1, 1)
    pkg.train_rectified_flow(m, x0, x1, epochs=2, batch_size=16, lr=1e-3)
    xt, x1t, t = torch.randn(2, 3, size, size, generator=gen), torch.randn(2, 3, size, size, generator=gen), torch.rand(2, generator=gen)
    m.eval()
    print("t =", t)
    report("test sequence", m, xt, x1t, t, arch, size, repeat=3)


if __name__ == "__main__":
    main()
