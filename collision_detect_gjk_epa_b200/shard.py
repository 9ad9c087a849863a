"""Pair sharding across the GPUs of one box.

Pairs are independent, so the batch splits by contiguous pair-index range with NO collective on
the data path: every rank (one process per GPU) holds the whole hull store (replicated; 3.84 GB at
cfg5) and its slice of the pair list.  ShardedEngine.run returns SHARD-LOCAL arrays (row k = pair
pair_range[0] + k); only the gathered contact list carries global pair numbers.  The only
exchange is optional and follows the kernels: an all-gather of the compacted contact list
(counts first, then padded blocks) so that every rank sees every contact.

One HOST process over several GPUs needs none of this: a context made by gjkepa_create_multi
(Engine(devices=[...])) shards inside the library and writes at the global pair index.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of pair indices owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_pairs), world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def balanced_ranges(cost, world: int):
    """Cost-balanced contiguous ranges: cost[i] ~ nA+nB of pair i (support scans dominate large
    hulls).  Returns a list of (lo, hi)."""
    import numpy as np
    c = np.cumsum(np.asarray(cost, dtype=np.float64))
    total = float(c[-1]) if len(c) else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, total * r / world, side="left")))
    cuts.append(len(c))
    cuts = [max(cuts[i], cuts[i - 1]) if i else 0 for i in range(len(cuts))]
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def compact_contacts(collision, depth, normal, pair_base: int = 0):
    """Indices (global), depth and normal of the colliding pairs of this shard (torch tensors)."""
    import torch
    idx = torch.nonzero(collision, as_tuple=False).flatten()
    return (idx + pair_base).to(torch.int64), depth[idx], normal[idx]


def gather_contacts(idx, depth, normal, group=None):
    """All-gather a ragged contact list: every rank receives the concatenation over ranks, in rank
    order (= ascending pair index when shards are contiguous ranges).  Works with NCCL (CUDA
    tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return idx, depth, normal
    world = dist.get_world_size(group)
    n_local = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    # one padded block per rank: [index as float64 bits | depth | normal(3)]
    block = torch.zeros((cap, 5), dtype=torch.float64, device=idx.device)
    k = idx.numel()
    if k:
        block[:k, 0] = idx.to(torch.int64).view(torch.float64) if idx.dtype == torch.int64 else idx.double()
        block[:k, 1] = depth
        block[:k, 2:5] = normal
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    parts = [blocks[r][:counts[r]] for r in range(world)]
    allb = torch.cat(parts, dim=0) if parts else block[:0]
    return allb[:, 0].contiguous().view(torch.int64), allb[:, 1].contiguous(), allb[:, 2:5].contiguous()


class ShardedEngine:
    """Runs this rank's slice of a batch on its GPU; `engine` is a collision_detect_gjk_epa_b200.Engine."""

    def __init__(self, engine, rank: int, world: int, group=None):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group

    def run(self, verts, hull_off, pair_a, pair_b, mode="epa", gather: bool = False, **kw):
        lo, hi = shard_range(int(pair_a.shape[0]), self.rank, self.world)
        out = self.engine.batch_device(verts, hull_off, pair_a[lo:hi].contiguous(), pair_b[lo:hi].contiguous(),
                                       mode=mode, **kw)
        out["pair_range"] = (lo, hi)
        if gather and "depth" in out:
            out["contacts"] = gather_contacts(*compact_contacts(out["collision"], out["depth"], out["normal"], lo),
                                              group=self.group)
        return out


def bind_to_gpu_numa(device_index: int):
    """One process per GPU: run this process (and so first-touch its pinned host buffers) on the CPU cores of the NUMA node
    the GPU hangs off.  With 8 ranks on a two-socket box, half of them would otherwise stream their H2D copies from the
    other socket's memory.  Best effort: returns (node, previous_affinity) or None when the topology cannot be read or the
    cpuset forbids it; `os.sched_setaffinity(0, previous_affinity)` undoes it."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        prev = os.sched_getaffinity(0)
        use = cpus & prev
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return node, prev
    except Exception:
        return None
