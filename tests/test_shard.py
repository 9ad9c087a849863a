"""Pair sharding: index math on one process, and the N>1 contact gather on CPU with
world_size = 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from collision_detect_gjk_epa_b200 import shard


@pytest.mark.parametrize("n,world", [(0, 4), (1, 4), (7, 2), (8, 8), (1_000_003, 8), (100_000_000, 8)])
def test_shard_ranges_partition_the_batch(n, world):
    r = [shard.shard_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1


def test_balanced_ranges_even_out_cost():
    rng = np.random.default_rng(0)
    cost = rng.integers(8, 129, size=100000)
    rr = shard.balanced_ranges(cost, 8)
    assert rr[0][0] == 0 and rr[-1][1] == len(cost) and all(rr[k][1] == rr[k + 1][0] for k in range(7))
    tot = [cost[a:b].sum() for a, b in rr]
    assert max(tot) / min(tot) < 1.01


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a deterministic fake narrow phase: pair i collides iff i % 3 == 0, depth = i/10, normal = (1,0,0)*sign
    lo, hi = shard.shard_range(n_pairs, rank, world)
    i = torch.arange(lo, hi)
    col = (i % 3 == 0).to(torch.uint8)
    depth = i.double() / 10.0
    normal = torch.stack([torch.ones_like(depth), torch.zeros_like(depth), -torch.ones_like(depth)], dim=1)
    idx, d, n = shard.gather_contacts(*shard.compact_contacts(col, depth, normal, lo))
    q.put((rank, idx.numpy(), d.numpy(), n.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [1000, 5])
def test_contact_gather_world_size_2_gloo(n_pairs):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(0, n_pairs, 3)
    for rank, idx, d, n in got:
        assert np.array_equal(idx, want)                       # rank order = ascending pair index
        assert np.allclose(d, want / 10.0) and np.all(n[:, 0] == 1.0) and np.all(n[:, 2] == -1.0)


def test_single_process_gather_is_identity():
    idx = torch.tensor([3, 9], dtype=torch.int64); d = torch.tensor([0.1, 0.2], dtype=torch.float64)
    n = torch.zeros((2, 3), dtype=torch.float64)
    a, b, c = shard.gather_contacts(idx, d, n)
    assert a is idx and b is d and c is n
