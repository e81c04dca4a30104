"""2-GPU check of the multi-GPU paths on real hardware (run under torchrun): sharded pair generation with the final NCCL
all-gather, and one data-parallel training step (replicas must stay identical, gradient = global mean)."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rectified_flow_vision_b200 as pkg
from rectified_flow_vision_b200 import dist as rdist
from rectified_flow_vision_b200.training import NativeTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(image_size=32, channel_mult=[1, 2], num_res_blocks=1)
torch.manual_seed(5)
m = pkg.RectifiedFlowModel(device=f"cuda:{local}", **kw)
m.eval()
# --- sharded pair generation, gathered ---
x0, x1 = rdist.generate_reflow_pairs_sharded(m, num_pairs=26, num_steps=3, seed=42, gather=True)
assert x0.shape == (26, 3, 32, 32) and x1.shape == x0.shape and x1.device.type == "cpu"
ref = m.sample(noise=x0.cuda(), num_steps=3).cpu()          # the whole job on this rank alone
err = float(((x1 - ref) ** 2).sum().sqrt() / (ref ** 2).sum().sqrt())
chk = torch.tensor([float(x1.double().sum())], device="cuda")
all_chk = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(all_chk, chk)
assert err < 2e-2, err
buf = torch.empty(26, 3, 32, 32).pin_memory()            # caller-owned pinned destination, re-used across gathers
for _ in range(2):
    _, x1b = rdist.generate_reflow_pairs_sharded(m, num_pairs=26, num_steps=3, seed=42, gather=True, out=buf)
    assert x1b is buf
    eb = float(((x1b - ref) ** 2).sum().sqrt() / (ref ** 2).sum().sqrt())
    rr = float(((x1b - x1) ** 2).sum().sqrt() / (x1 ** 2).sum().sqrt())   # run to run: summation order of the GroupNorm statistics
    if rank == 0:
        print(f"out= gather: rel-L2 vs single-rank {eb:.2e}, vs the first gather {rr:.2e}")
    assert eb < 2e-2 and rr < 2e-2, (eb, rr)
_, x1c = rdist.generate_reflow_pairs_sharded(m, num_pairs=25, num_steps=3, seed=42, gather=True, out=buf[:25])   # uneven shards
ec = float(((x1c - ref[:25]) ** 2).sum().sqrt() / (ref[:25] ** 2).sum().sqrt())
assert x1c.shape[0] == 25 and ec < 2e-2, ec
assert all(abs(float(c) - float(all_chk[0])) < 1e-3 for c in all_chk), all_chk       # every rank holds the same gathered tensor
# --- one data-parallel training step ---
g = torch.Generator().manual_seed(100)
gx0, gx1, gt = torch.randn(8, 3, 32, 32, generator=g), torch.randn(8, 3, 32, 32, generator=g), torch.rand(8, generator=g)
lo, hi = rdist.shard_bounds(8, rank, world)
tr = NativeTrainer(m, lr=1e-3, micro_batch=4)
loss = tr.step(gx0[lo:hi].cuda(), gx1[lo:hi].cuda(), gt[lo:hi].cuda())
psum = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum().reshape(1)
all_p = [torch.zeros_like(psum) for _ in range(world)]
dist.all_gather(all_p, psum)
assert all(float(p) == float(all_p[0]) for p in all_p), all_p                          # replicas bit-identical after the step
# single-process reference of the same global step
if rank == 0:
    torch.manual_seed(5)
    m1 = pkg.RectifiedFlowModel(device=f"cuda:{local}", **kw)
    m1.eval()
    NativeTrainer(m1, lr=1e-3, micro_batch=8, process_group=None)
    import torch.distributed as d2
    # emulate world=1: call the engine directly (the trainer would all-reduce in this process group)
    eng = m1.velocity_net.train_engine(32, f"cuda:{local}", 8)
    eng.zero_grad()
    eng.train_accumulate(gx0.cuda(), gx1.cuda(), gt.cuda(), 0.0, 1)
    eng.optimizer_step(1e-3, 1)
    torch.manual_seed(5)
    p0 = [q.detach().clone() for q in pkg.RectifiedFlowModel(device=f"cuda:{local}", **kw).parameters()]
    num = sum(float(((a.detach() - b.detach()) ** 2).sum()) for a, b in zip(m.parameters(), m1.parameters()))
    den = sum(float(((b.detach() - c) ** 2).sum()) for b, c in zip(m1.parameters(), p0))
    d = (num / den) ** 0.5
    print(f"pair generation rel-L2 vs single-rank {err:.2e}; DP step vs single-process step: update rel-L2 {d:.2e}; loss {float(loss):.4f}")
    # Adam's first update is +-lr per element: where a near-zero gradient changes sign under bf16 noise the parameter moves by 2*lr,
    # so the comparison is on the whole update vector, not element-wise
    assert d < 0.25, d
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}: ok")
