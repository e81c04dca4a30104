"""Per-tensor gradient comparison: native training step vs the CPU oracle (run on the GPU box)."""
import argparse
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="small32")
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--mb", type=int, default=4)
    a = ap.parse_args()
    import rectified_flow_vision_b200 as pkg
    from oracle import train_oracle as T
    m = util.seeded_model(a.case, device="cuda:0", cls=pkg.RectifiedFlowModel)
    kw = util.manifest()["cases"][a.case]["kwargs"]
    arch = dict(model_channels=kw.get("model_channels", 64), channel_mult=tuple(kw.get("channel_mult", [1, 2, 4])),
                num_res_blocks=kw.get("num_res_blocks", 2))
    gen = torch.Generator().manual_seed(7)
    S = kw["image_size"]
    x0, x1, t = torch.randn(a.batch, 3, S, S, generator=gen), torch.randn(a.batch, 3, S, S, generator=gen), torch.rand(a.batch, generator=gen)
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    loss_ref, grads = T.loss_and_grads(P, x0, x1, t, **arch)
    eng = m.velocity_net.train_engine(S, "cuda:0", micro_batch=a.mb)
    eng.zero_grad()
    loss = float(eng.train_accumulate(x0.cuda(), x1.cuda(), t.cuda(), dropout_p=0.0, seed=3).item())
    torch.cuda.synchronize()
    print(f"loss {loss:.6f} ref {loss_ref:.6f}")
    print(f"{'tensor':<52}{'|g_ref|':>12}{'|g|':>12}{'rel-L2':>12}")
    for k, gr in grads.items():
        g = eng.get_grad(k, gr.numel()).cpu().numpy().reshape(gr.shape)
        err = util.rel_l2(g, gr.numpy())
        flag = "" if err < 0.1 else "  <<<<"
        print(f"{k[13:]:<52}{float(gr.norm()):>12.4e}{float(np.sqrt((g.astype(np.float64)**2).sum())):>12.4e}{err:>12.3e}{flag}")


if __name__ == "__main__":
    main()
