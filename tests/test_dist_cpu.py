"""world_size-2 gloo test of the batch-sharding logic (the host side of the multi-GPU path).  The per-shard compute
is a stand-in function here: on a CPU box there is no CUDA path to call (and no fallback by design)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rectified_flow_vision_b200 import dist as rdist


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 65536, 1001):
        for w in (1, 2, 3, 4, 8):
            spans = [rdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_seeded_noise_is_rank_independent():
    a = rdist.seeded_noise(5, 3, 8, 42)
    b = rdist.seeded_noise(5, 3, 8, 42)
    assert torch.equal(a, b) and a.shape == (5, 3, 8, 8)
    assert torch.equal(a, torch.randn(5, 3, 8, 8, generator=torch.Generator().manual_seed(42)))


def _fake_integrate(x):
    return x * 2.0 + 1.0


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = rdist.seeded_noise(n, 3, 4, 7)
    full = rdist.sharded_map(_fake_integrate, rows, gather=True)
    part = rdist.sharded_map(_fake_integrate, rows, gather=False)
    buf = torch.full_like(rows, -5.0)                     # caller-owned destination, re-used across two gathers
    got = rdist.sharded_map(_fake_integrate, rows, gather=True, out=buf)
    assert got is buf
    rdist.sharded_map(_fake_integrate, rows, gather=True, out=buf)
    even = rows[: (n // world) * world]                   # equal shards: the collective lands in `out` without a re-pack
    buf2 = torch.empty_like(even)
    assert rdist.sharded_map(_fake_integrate, even, gather=True, out=buf2) is buf2
    torch.save({"full": full, "part": part, "out": buf, "out_even": buf2}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_map_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n = 7  # uneven split: 4 + 3
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    rows = rdist.seeded_noise(n, 3, 4, 7)
    want = _fake_integrate(rows)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["full"], want) and torch.equal(r1["full"], want)
    assert torch.equal(r0["part"], want[:4]) and torch.equal(r1["part"], want[4:])
    assert torch.equal(r0["out"], want) and torch.equal(r1["out"], want)
    assert torch.equal(r0["out_even"], want[:6]) and torch.equal(r1["out_even"], want[:6])


def test_single_process_passthrough():
    rows = rdist.seeded_noise(3, 3, 4, 1)
    assert torch.equal(rdist.sharded_map(_fake_integrate, rows), _fake_integrate(rows))
    buf = torch.empty_like(rows)
    assert rdist.sharded_map(_fake_integrate, rows, out=buf) is buf and torch.equal(buf, _fake_integrate(rows))


# ---- data-parallel training step: host-side sequence zero_grad -> accumulate -> all-reduce(SUM) -> step(1/world) ----
class _StubEngine:
    """Stands in for the native handle on a CPU box: gradient = mean of the shard, 'optimizer' = SGD on one weight."""

    def __init__(self):
        self.g = torch.zeros(3)
        self.w = torch.zeros(3)
        self.calls = []

    def zero_grad(self):
        self.g.zero_()
        self.calls.append("zero")

    def reset_optimizer(self):
        self.calls.append("reset")

    def train_accumulate(self, x0, x1, t, dropout_p=0.0, seed=0):
        self.g += (x1 - x0).mean(dim=0)
        self.calls.append(("acc", dropout_p, seed))
        return ((x1 - x0) ** 2).mean()

    def grad_buffer(self):
        return self.g

    def optimizer_step(self, lr, step, b1, b2, eps, wd, max_norm, grad_scale=1.0):
        self.calls.append(("step", step, grad_scale))
        self.w -= lr * self.g * grad_scale
        return (self.g * grad_scale).norm()


class _StubNet:
    dropout_p = 0.1

    def __init__(self):
        self.eng = _StubEngine()

    def train_engine(self, size, device, micro_batch=None):
        return self.eng


class _StubModel:
    device = "cpu"
    training = True

    def __init__(self):
        self.velocity_net = _StubNet()


def _train_worker(rank, world, port, out_dir):
    from rectified_flow_vision_b200.training import NativeTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _StubModel()
    tr = NativeTrainer(m, lr=0.5)
    data = torch.arange(12, dtype=torch.float32).view(4, 3)       # the global batch: 4 rows
    lo, hi = rdist.shard_bounds(4, rank, world)
    x1 = data[lo:hi]
    tr.step(torch.zeros_like(x1), x1, torch.zeros(hi - lo))
    torch.save({"w": m.velocity_net.eng.w, "calls": m.velocity_net.eng.calls}, os.path.join(out_dir, f"t{rank}.pt"))
    dist.destroy_process_group()


def test_data_parallel_trainer_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    want = -0.5 * torch.arange(12, dtype=torch.float32).view(4, 3).mean(dim=0)   # one SGD step on the GLOBAL mean gradient
    assert torch.allclose(r0["w"], want) and torch.allclose(r1["w"], want)       # identical replicas after the step
    # a new trainer zeroes the Adam moments first; the dropout seed is the step counter with the rank in the high half
    assert r0["calls"][:2] == ["reset", "zero"] and r0["calls"][2] == ("acc", 0.1, 1) and r0["calls"][3] == ("step", 1, 0.5)
    assert r1["calls"][2] == ("acc", 0.1, 1 | (1 << 32))
