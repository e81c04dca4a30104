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
    torch.save({"full": full, "part": part}, os.path.join(out_dir, f"r{rank}.pt"))
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


def test_single_process_passthrough():
    rows = rdist.seeded_noise(3, 3, 4, 1)
    assert torch.equal(rdist.sharded_map(_fake_integrate, rows), _fake_integrate(rows))
