"""BASELINE.json configs[4], literally: a 100M-pair broad-phase output batch over a shared store of 16-vertex hulls,
sharded by contiguous pair-index range across the GPUs of one box, contact list (pair index, depth, normal) gathered
and copied to the host.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tools/cfg5_sharded.py --bodies 10000000 --pairs 100000000

Every rank builds the same store and pair list (same seed), keeps the store resident (replicated, 3.84 GB), runs its
range in chunks, compacts its contacts and joins the all-gather (collision_detect_gjk_epa_b200/shard.py; NCCL).  Timed
region (device events + barrier, max over ranks): kernels of all chunks + compaction + all-gather + device->host copy
of the gathered list on rank 0.  Afterwards rank 0 recomputes the WHOLE batch alone and compares the contact list bit
for bit (sharding must not change a single flag, depth or normal).  Prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=10_000_000)
    ap.add_argument("--pairs", type=int, default=100_000_000)
    ap.add_argument("--chunk", type=int, default=12_500_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from collision_detect_gjk_epa_b200 import Engine, shard, workloads as W

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t_gen = time.perf_counter()
    b = W.make_shared_store(args.bodies, args.pairs, n_vert=16, seed=2026)
    t_gen = time.perf_counter() - t_gen
    eng = Engine(local)
    eng.set_max_hull_vertices(16)
    verts = torch.from_numpy(b.verts).to(dev)
    hull_off = torch.from_numpy(b.hull_off).to(dev)
    lo, hi = shard.shard_range(args.pairs, rank, world)
    pa = torch.from_numpy(b.pair_a[lo:hi]).to(dev)
    pb = torch.from_numpy(b.pair_b[lo:hi]).to(dev)
    n = hi - lo
    col = torch.empty(n, dtype=torch.uint8, device=dev)
    dep = torch.empty(n, dtype=torch.float64, device=dev)
    nrm = torch.empty((n, 3), dtype=torch.float64, device=dev)

    def run_range(pa_, pb_, col_, dep_, nrm_):
        for c0 in range(0, pa_.shape[0], args.chunk):
            c1 = min(pa_.shape[0], c0 + args.chunk)
            eng.batch_device(verts, hull_off, pa_[c0:c1], pb_[c0:c1], mode="epa", want_status=False,
                             out={"collision": col_[c0:c1], "depth": dep_[c0:c1], "normal": nrm_[c0:c1]})

    host = {}

    def to_host(name, t):                                # pinned buffers, allocated once
        if name not in host or host[name].shape[0] < t.shape[0]:
            host[name] = torch.empty((int(t.shape[0] * 1.05) + 16,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
        h = host[name][:t.shape[0]]
        h.copy_(t, non_blocking=True)
        return h

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step():
        ev[0].record()
        run_range(pa, pb, col, dep, nrm)
        ev[1].record()
        mine = shard.compact_contacts(col, dep, nrm, lo)
        ev[2].record()
        idx, d, nv = shard.gather_contacts(*mine)
        ev[3].record()
        out = None
        if rank == 0:
            out = (to_host("idx", idx), to_host("d", d), to_host("n", nv))      # the gathered list on the host
        ev[4].record()
        torch.cuda.synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()                                               # warm-up (workspace allocation, NCCL channels)
    best = None
    for _ in range(args.reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        got = step()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall] + [ev[i].elapsed_time(ev[i + 1]) * 1e-3 for i in range(4)],
                         dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.tolist()
        best = t if best is None or t[1] < best[1] else best
    line = None
    if rank == 0:
        idx, d, nv = got
        line = {"workload": f"cfg5: {args.pairs} pairs over a shared store of {args.bodies} hull(16) bodies, GJK+EPA, contact "
                            "list gathered to the host", "n_gpus": world, "pairs": args.pairs, "contacts": int(idx.numel()),
                "seconds_device_events_max_over_ranks": best[0], "seconds_wall_max_over_ranks": best[1],
                "phases_s_max_over_ranks": dict(zip(["kernels", "compaction", "all_gather", "copy_to_host"], best[2:])),
                "pairs_per_s": args.pairs / best[1], "contacts_per_s": int(idx.numel()) / best[1],
                "contact_list_bytes_to_host": int(idx.numel()) * 40, "hull_store_bytes_per_gpu": int(b.verts.nbytes),
                "host_generation_s": t_gen, "sharding": f"contiguous pair-index ranges x{world}, store replicated"}
        if not args.no_verify and world > 1:
            # rank 0 alone over the whole batch: the gathered list must be identical
            del pa, pb, col, dep, nrm
            fa = torch.from_numpy(b.pair_a).to(dev); fb = torch.from_numpy(b.pair_b).to(dev)
            fc = torch.empty(args.pairs, dtype=torch.uint8, device=dev)
            fd = torch.empty(args.pairs, dtype=torch.float64, device=dev)
            fn = torch.empty((args.pairs, 3), dtype=torch.float64, device=dev)
            run_range(fa, fb, fc, fd, fn)
            i1, d1, n1 = shard.compact_contacts(fc, fd, fn, 0)
            idx, d, nv = idx.clone(), d.clone(), nv.clone()
            line["identical_to_single_gpu"] = bool(torch.equal(i1.cpu(), idx) and torch.equal(d1.cpu(), d) and torch.equal(n1.cpu(), nv))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
