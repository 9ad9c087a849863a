#!/usr/bin/env python3
"""bench.py -- the GJK(+EPA) hot path on synthetic convex-hull pairs (BASELINE.json / SURVEY.md §8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg4deep|cfg5]
    python bench.py --impl reference ...      # the CPU path (oracle port, OpenMP over pairs)

A "step" is one pass of the hot path over one batch.  N=1 workload: BASELINE.json configs[1]
(1M random pairs of 8-vertex boxes, GJK intersection test only).  For N>1 (torchrun, one rank per
GPU) the pairs shard by index: every rank runs its own 1M-pair shard, no data-path collective
("weak" scaling).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gjk_pair_queries_per_sec"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary GJK+EPA measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def make_workload(name: str, n_pairs: int, seed: int):
    from collision_detect_gjk_epa_b200 import workloads as W
    if name == "cfg2":
        return W.make_pairs("box8", n_pairs, 1.0, seed=seed), "gjk", \
            "cfg2: 1M random pairs of 8-vertex boxes (random pose, offset N(0,1)), GJK intersection test only"
    if name == "cfg3":
        return W.make_pairs("hull", n_pairs, 0.85, seed=seed, n_vert=32), "epa", \
            "cfg3: 1M pairs of 32-vertex hulls, ~52% overlapping, GJK+EPA depth and normal"
    if name == "cfg4":
        return W.make_mixed(n_pairs, 0.25, seed=seed), "epa", "cfg4: mixed 4-64 vertex hulls, offset N(0,0.25), GJK+EPA"
    if name == "cfg4deep":
        return W.make_pairs("hull", n_pairs, 0.05, seed=seed, n_vert=64), "epa", "cfg4 deep slice: 64-vertex hulls, offset N(0,0.05)"
    if name == "cfg5":
        return W.make_shared_store(max(1000, n_pairs // 10), n_pairs, n_vert=16, seed=seed), "epa", \
            "cfg5: pairs over a shared store of 16-vertex hulls (10 pairs per body), GJK+EPA"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.live = threading.Event()       # samples are taken only while a timed region runs
        self.live.set()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            if not self.live.is_set():
                time.sleep(self.period)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def cpu_rate(batch, mode, sample_pairs, threads=0, repeats=1):
    """Times the CPU path (oracle port: C restatement of GCLIB_GJKEPA.f90, OpenMP over pairs)."""
    from oracle import orc
    n = min(sample_pairs, batch.n_pairs)
    if threads <= 0:                      # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
    omode = {"gjk": orc.MODE_GJK, "epa": orc.MODE_EPA, "full": orc.MODE_FULL}[mode]
    pa, pb = batch.pair_a[:n], batch.pair_b[:n]
    best, used, hits = None, 1, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        o = orc.batch(batch.verts, batch.hull_off, pa, pb, mode=omode, n_threads=threads)
        dt = time.perf_counter() - t0
        used, hits = o["threads"], int(o["collision"].sum())
        best = dt if best is None else min(best, dt)
    return n / best, used, n, hits, best


def static_config(desc, n, mode, world, bytes_per_pair, store_mb):
    """The part of `config` both arms share (the reference arm runs the same pairs on the host cores)."""
    return {"workload": desc, "pairs_per_gpu_per_step": n, "mode": mode, "sharding": f"pairs by index x{world}",
            "l2": "inputs larger than L2 (hull store %.0f MB per GPU)" % store_mb,
            "algorithmic_bytes_per_pair": bytes_per_pair}


def run_reference(args, rank, world):
    """The CPU path on the box's host cores: the C restatement of GCLIB_GJKEPA.f90 (oracle/), OpenMP over pairs, every
    core, the SAME batch as the GPU arm's rank 0 (same seed, same number of pairs per step)."""
    if rank != 0:
        return
    batch, mode, desc = make_workload(args.workload, args.pairs, seed=12345)
    out_bytes = {"gjk": 1, "epa": 33, "full": 109}[mode]
    n = batch.n_pairs
    for _ in range(args.warmup):
        cpu_rate(batch, mode, min(n, 50_000))
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        rate, used, ns, hits, dt = cpu_rate(batch, mode, n)
        n_done += ns
    total = time.perf_counter() - t0
    value = n_done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(desc, n, mode, world, batch.algorithmic_bytes_per_pair(out_bytes), batch.verts.nbytes / 1e6),
        "colliding_fraction": hits / n,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"all {n} pairs of rank 0's batch per step; C restatement of "
                                   "GCLIB_GJKEPA.f90 (not gfortran: no Fortran compiler in the image), "
                                   "OpenMP parallel-for over pairs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_one_process(args, world, batch, h_in, mode, pinned_empty, W, dev):
    """Rank 0 alone drives all N GPUs of the box through ONE multi-device context (gjkepa_create_multi), the other ranks
    idle: the call a single Fortran host process makes.  Pairs shard by contiguous index range inside the library, every
    GPU copies its own results into the caller's arrays, no collective.  Two workloads: the cfg2 step with everything
    crossing PCIe (N x the per-GPU batch, so the per-GPU work is the same as in the one-rank-per-GPU run above), and
    BASELINE configs[4] scaled to N GPUs: 12.5M pairs per GPU over a shared store of 16-vertex hulls resident on every
    GPU (N = 8: 100M pairs over 10M bodies), GJK + EPA, the contact list (pair index, depth, normal) gathered to the host."""
    import ctypes as C
    import torch
    from collision_detect_gjk_epa_b200 import Engine
    ndev = min(world, torch.cuda.device_count())
    eng = Engine(devices=list(range(ndev)))
    res = {"devices": ndev, "api": "gjkepa_create_multi + host-buffer batch calls (C ABI), one host process"}

    def timed(fn, steps):
        fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        return (time.perf_counter() - t0) / steps

    def phases():
        ms = (C.c_double * 2)()
        eng._L.gjkepa_host_phase_ms(eng._h, ms)
        return {"shards_ms": ms[0], "contact_rows_to_host_ms": ms[1]}

    # (a) cfg2, copy everything: the per-GPU batch tiled N times
    n1 = batch.n_pairs
    nh1 = int(batch.hull_off.shape[0] - 1)
    nv1 = int(batch.hull_off[-1])
    verts = pinned_empty((ndev * batch.verts.shape[0],), torch.float64)
    off = pinned_empty((ndev * nh1 + 1,), torch.int64)
    pa = pinned_empty((ndev * n1,), torch.int32)
    pb = pinned_empty((ndev * n1,), torch.int32)
    for r in range(ndev):
        verts[r * batch.verts.shape[0]:(r + 1) * batch.verts.shape[0]] = batch.verts
        off[r * nh1:(r + 1) * nh1 + 1] = batch.hull_off + r * nv1
        pa[r * n1:(r + 1) * n1] = batch.pair_a + r * nh1
        pb[r * n1:(r + 1) * n1] = batch.pair_b + r * nh1
    flags = pinned_empty((ndev * n1,), torch.uint8)
    eng.set_max_hull_vertices(int(np.diff(batch.hull_off).max()))
    steps = max(3, min(args.steps, 10))
    dt = timed(lambda: eng.batch_host(verts, off, pa, pb, mode=mode, out={"collision": flags}), steps)
    assert np.array_equal(flags[:n1], flags[(ndev - 1) * n1:]), "shards of a tiled batch must agree"
    res["cfg2_copy_all"] = {"pairs_per_step": ndev * n1, "ms_per_step": dt * 1e3, "value": ndev * n1 / dt, "unit": UNIT,
                            "h2d_bytes_per_step": int(verts.nbytes + off.nbytes + pa.nbytes + pb.nbytes), "d2h_bytes_per_step": int(flags.nbytes),
                            "h2d_gbs_total": (verts.nbytes + off.nbytes + pa.nbytes + pb.nbytes) / dt / 1e9}
    eng.upload_store(verts, off)
    dt = timed(lambda: eng.batch_resident(pa, pb, mode=mode, out={"collision": flags}), steps)
    res["cfg2_resident_store"] = {"pairs_per_step": ndev * n1, "ms_per_step": dt * 1e3, "value": ndev * n1 / dt, "unit": UNIT,
                                  "h2d_bytes_per_step": int(pa.nbytes + pb.nbytes), "d2h_bytes_per_step": int(flags.nbytes)}
    del verts, off, pa, pb, flags

    # (b) BASELINE configs[4]: shared store of 16-vertex hulls, 10 pairs per body, contact list to the host
    per_gpu = int(os.environ.get("GJKEPA_BENCH_CFG5_PAIRS_PER_GPU", 12_500_000))
    n5, nb5 = per_gpu * ndev, max(1000, per_gpu * ndev // 10)
    t0 = time.perf_counter()
    v5, o5, a5, b5 = W.make_shared_store_gpu(dev, nb5, n5, n_vert=16, seed=4242)
    v5, o5, a5, b5 = v5.numpy(), o5.numpy(), a5.numpy(), b5.numpy()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.upload_store(v5, o5)
    up_s = time.perf_counter() - t0
    out5 = {"collision": pinned_empty((n5,), torch.uint8), "contact_index": pinned_empty((n5,), torch.int32),
            "depth": pinned_empty((n5,), torch.float64), "normal": pinned_empty((n5, 3), torch.float64)}
    dt = timed(lambda: eng.batch_resident(a5, b5, mode="epa", out=out5, compact=True), 3)
    nc = int(out5["n_contacts"])
    ph = phases()
    idx = out5["contact_index"][:nc]
    ok = bool(nc > 0 and np.all(np.diff(idx.astype(np.int64)) > 0) and np.all(out5["collision"][idx] == 1)
              and int(out5["collision"].sum(dtype=np.int64)) == nc and np.all(out5["depth"][:nc] >= 0.0)
              and np.allclose(np.linalg.norm(out5["normal"][:min(nc, 1_000_000)], axis=1), 1.0, atol=1e-9))
    res["cfg5"] = {"workload": f"BASELINE configs[4] at {ndev} GPU(s): {n5} pairs over a shared store of {nb5} 16-vertex hulls "
                               f"({v5.nbytes / 1e9:.2f} GB, resident on every GPU), GJK+EPA, contact list (pair index, depth, normal) to the host",
                   "pairs_per_step": n5, "contacts": nc, "ms_per_step": dt * 1e3, "pairs_per_s": n5 / dt, "contacts_per_s": nc / dt,
                   "h2d_bytes_per_step": int(8 * n5), "d2h_bytes_per_step": int(n5 + 36 * nc), "phases": ph,
                   "store_upload_s": up_s, "workload_generation_s": gen_s,
                   "contact_list_checks_ok": ok,
                   "algorithmic_gbs": (24 * 32 + 8 + 33) * n5 / dt / 1e9}
    dtf = timed(lambda: eng.batch_resident(a5, b5, mode="gjk", out={"collision": out5["collision"]}), 3)
    res["cfg5"]["gjk_only"] = {"ms_per_step": dtf * 1e3, "pairs_per_s": n5 / dtf}
    # The same pairs in the order a broad phase emits them -- grouped by their first body -- instead of the random order above:
    # consecutive pairs then share a hull and their partners are lattice neighbours, so the gathers from the 16-vertex store hit
    # L2 instead of DRAM.  Same contacts (checked by count and by the multiset of pair numbers' flags).
    ta = torch.from_numpy(a5).to(dev)
    order = torch.argsort(ta, stable=True)
    a5s = torch.from_numpy(a5); b5s = torch.from_numpy(b5)
    sa_ = ta[order].cpu(); sb_ = torch.from_numpy(b5).to(dev)[order].cpu()
    del ta, order
    torch.cuda.empty_cache()
    a5s.copy_(sa_); b5s.copy_(sb_)                        # in place: a5 / b5 are views of the pinned tensors
    dts = timed(lambda: eng.batch_resident(a5, b5, mode="epa", out=out5, compact=True), 3)
    res["cfg5"]["broad_phase_order"] = {"what": "the same pair list sorted by first body (the order a broad phase emits)",
                                        "ms_per_step": dts * 1e3, "pairs_per_s": n5 / dts, "contacts_per_s": int(out5["n_contacts"]) / dts,
                                        "contacts": int(out5["n_contacts"]), "same_contact_count": int(out5["n_contacts"]) == nc, "phases": phases()}
    eng.close()
    return res


def emit(line: dict):
    """The one JSON line goes to the real stdout; everything else this process or its libraries print
    (NCCL's version banner, warnings) was diverted to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # fd 1 -> stderr for libraries that print to stdout
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from collision_detect_gjk_epa_b200 import Engine
    from collision_detect_gjk_epa_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GJK/EPA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from collision_detect_gjk_epa_b200 import shard as _shard
    numa = _shard.bind_to_gpu_numa(local_rank)       # pinned buffers on the GPU's own NUMA node (undone before the CPU baseline)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # ranks that sit out a leg wait on the CPU, not in an NCCL kernel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    keep = []

    def pin(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(tt)
        return tt.numpy()

    def pinned_empty(shape, dtype):
        tt = torch.empty(shape, dtype=dtype, pin_memory=True)
        keep.append(tt)
        return tt.numpy()

    batch, mode, desc = make_workload(args.workload, args.pairs, seed=12345 + rank)
    eng = Engine(local_rank, lanes=args.lanes)
    max_verts = int(np.diff(batch.hull_off).max())
    eng.set_max_hull_vertices(max_verts)      # the caller knows its hull sizes
    n = batch.n_pairs
    out_bytes = {"gjk": 1, "epa": 33, "full": 109}[mode]
    bytes_per_pair = batch.algorithmic_bytes_per_pair(out_bytes)

    # ---- device-resident inputs: `value` ----
    d_in = [torch.from_numpy(x).to(dev) for x in (batch.verts, batch.hull_off, batch.pair_a, batch.pair_b)]
    out = None

    def step_device():
        nonlocal out
        out = eng.batch_device(*d_in, mode=mode, out=out, want_status=False)

    warm = max(3, args.warmup)
    for _ in range(warm):
        step_device()
    barrier()
    eng.launch_count(reset=True)
    kern = {"hull_prep": 0.0, "gjk": 0.0, "epa": 0.0}
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    sampler.live.clear()
    ms_total = reduce_max(e0.elapsed_time(e1))
    launches = eng.launch_count()
    # per-kernel durations: a few more (untimed) profiled steps give a per-launch sample of each kernel
    ksamples = []
    eng.set_profiling(True)
    for _ in range(5):
        step_device()
        ksamples.append(eng.kernel_ms())
    for k in kern:
        kern[k] = float(np.median([s_[k] for s_ in ksamples]))
    eng.set_profiling(False)
    hits = int(out["collision"].sum().item())
    pairs_all, hits_all = reduce_sum([float(n), float(hits)])
    value = pairs_all * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region): `e2e` ----
    h_in = [pin(x) for x in (batch.verts, batch.hull_off, batch.pair_a, batch.pair_b)]
    okeys = {"gjk": ["collision"], "epa": ["collision", "depth", "normal"],
             "full": ["collision", "depth", "normal", "colli_type", "nearest", "point"]}[mode]
    shapes = {"collision": ((n,), torch.uint8), "depth": ((n,), torch.float64), "normal": ((n, 3), torch.float64),
              "colli_type": ((n,), torch.int32), "nearest": ((n, 6), torch.float64), "point": ((n, 3), torch.float64)}
    h_out = {k: pinned_empty(*shapes[k]) for k in okeys}
    h2d = int(sum(x.nbytes for x in h_in))
    d2h = int(sum(v.nbytes for v in h_out.values()))
    eng._L.gjkepa_use_own_stream(eng._h)
    e2e_steps = max(3, min(args.steps, 20))

    def timed_host(fn, steps, everyone=True):
        """steps calls of fn between two barriers; wall clock, max over the ranks taking part"""
        for _ in range(2):
            fn()
        if everyone:
            barrier()
        else:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reduce_max(dt) if everyone else dt

    sampler.live.set()
    dt_copy = timed_host(lambda: eng.batch_host(*h_in, mode=mode, out=h_out), e2e_steps)
    e2e_value = pairs_all * e2e_steps / dt_copy
    # the same call with the hull store resident on the device (gjkepa_store_upload once, as a simulation uploads its
    # geometry; per step only the pair list goes in and the results come out)
    eng.upload_store(h_in[0], h_in[1])
    dt_res = timed_host(lambda: eng.batch_resident(h_in[2], h_in[3], mode=mode, out=h_out), e2e_steps)
    clocks = sampler.stop()      # sampled through the timed regions above (device-resident steps, then end-to-end steps)
    e2e_res = {"value": pairs_all * e2e_steps / dt_res, "unit": UNIT, "ms_per_step": dt_res / e2e_steps * 1e3,
               "h2d_bytes_per_step": int(h_in[2].nbytes + h_in[3].nbytes), "d2h_bytes_per_step": d2h,
               "api": "gjkepa_store_upload once + gjkepa_batch_resident_f64 per step (C ABI, pinned host buffers)"}
    e2e_solo = None
    if world > 1:
        # the same copy-everything step with rank 0 alone on the box: what one GPU reaches when the other N-1 do not
        # share the host's memory system with it (the per-GPU rate of the concurrent run, over this, is the e2e efficiency)
        if rank == 0:
            dts = timed_host(lambda: eng.batch_host(*h_in, mode=mode, out=h_out), e2e_steps, everyone=False)
            e2e_solo = n * e2e_steps / dts
        host_barrier()

    # ---- plain pinned copies, all ranks at once: attributes the end-to-end number to the host/PCIe side ----
    def copy_rate(h2d_dir: bool):
        nb = 256 << 20
        hb = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        db = torch.empty(nb, dtype=torch.uint8, device=dev)
        src, dst = (hb, db) if h2d_dir else (db, hb)
        dst.copy_(src, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(8):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        gbs = 8 * nb / (c0.elapsed_time(c1) * 1e-3) / 1e9
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = gbs
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [round(float(v), 2) for v in t.tolist()]
    pcie = {"h2d_gbs_per_gpu_all_ranks_at_once": copy_rate(True), "d2h_gbs_per_gpu_all_ranks_at_once": copy_rate(False),
            "what": "8 x 256 MiB cudaMemcpyAsync from/to pinned host memory per rank, all ranks started together (CUDA events)"}

    # ---- GJK + EPA: contacts/s on cfg3 (1M pairs of 32-vertex hulls per GPU), device-resident and end to end ----
    gjk_epa = None
    if not args.no_extra and args.workload == "cfg2":
        b3, m3, d3 = make_workload("cfg3", args.pairs, seed=777 + rank)
        t3 = [torch.from_numpy(x).to(dev) for x in (b3.verts, b3.hull_off, b3.pair_a, b3.pair_b)]
        o3 = None
        eng.set_max_hull_vertices(32)
        for _ in range(3):
            o3 = eng.batch_device(*t3, mode=m3, out=o3, want_status=False)
        barrier()
        k3 = max(3, min(args.steps, 20))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k3):
            o3 = eng.batch_device(*t3, mode=m3, out=o3, want_status=False)
        f1.record()
        barrier()
        ms3 = reduce_max(f0.elapsed_time(f1)) / k3
        eng.set_profiling(True)
        ks3 = []
        for _ in range(3):
            o3 = eng.batch_device(*t3, mode=m3, out=o3, want_status=False)
            ks3.append(eng.kernel_ms())
        eng.set_profiling(False)
        h3 = int(o3["collision"].sum().item())
        p3_all, h3_all = reduce_sum([float(b3.n_pairs), float(h3)])
        # end to end: store resident, pair list in, compact contact rows (index, depth, normal) out
        eng._L.gjkepa_use_own_stream(eng._h)
        eng.upload_store(b3.verts, b3.hull_off)
        pa3, pb3 = pin(b3.pair_a), pin(b3.pair_b)
        n3 = b3.n_pairs
        c_out = {"collision": pinned_empty((n3,), torch.uint8), "contact_index": pinned_empty((n3,), torch.int32),
                 "depth": pinned_empty((n3,), torch.float64), "normal": pinned_empty((n3, 3), torch.float64)}
        dt3 = timed_host(lambda: eng.batch_resident(pa3, pb3, mode="epa", out=c_out, compact=True), k3)
        assert c_out["n_contacts"] == h3, (c_out["n_contacts"], h3)
        gjk_epa = {"workload": d3, "pairs_per_gpu_per_step": n3, "ms_per_step": ms3,
                   "pairs_per_s": p3_all / ms3 * 1e3, "contacts_per_s": h3_all / ms3 * 1e3, "contacts_per_step_all_gpus": h3_all,
                   "kernel_ms": {k: float(np.median([q[k] for q in ks3])) for k in ("hull_prep", "gjk", "epa")},
                   "algorithmic_gbs": b3.algorithmic_bytes_per_pair(33) * p3_all / (ms3 * 1e-3) / 1e9,
                   "timing": "device-resident inputs, CUDA events, max over ranks",
                   "e2e": {"contacts_per_s": h3_all * k3 / dt3, "pairs_per_s": p3_all * k3 / dt3, "ms_per_step": dt3 / k3 * 1e3,
                           "h2d_bytes_per_step": int(8 * n3), "d2h_bytes_per_step": int(n3 + 36 * h3),
                           "api": "gjkepa_batch_resident_f64 | GJKEPA_MODE_COMPACT (C ABI, pinned host buffers): pair list in, "
                                  "flags + one row (pair index, depth, normal) per contact out"}}
        del t3, o3
        eng.set_max_hull_vertices(max_verts)

    # ---- one HOST process over all N GPUs (gjkepa_create_multi): what a Fortran host calling the C ABI gets ----
    one_process = None
    if not args.no_extra and args.workload == "cfg2":
        host_barrier()
        if rank == 0:
            one_process = run_one_process(args, world, batch, h_in, mode, pinned_empty, W, dev)
        host_barrier()

    if rank != 0:
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "ncu_traffic.json"          # dram bytes per launch from the committed ncu --set full captures
    peaks, peak_src = None, "fallback"
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text()); peak_src = "measured"
    peak_gbs = float(peaks["hbm_gbs"]) if peaks else 6650.0
    dom = "epa" if (mode != "gjk" and kern["epa"] > kern["gjk"]) else "gjk"
    if tfile.exists():
        try:
            ent = json.loads(tfile.read_text()).get(args.workload, {})
            traffic = ent.get({"gjk": "gjk_kernel", "epa": "epa_kernel"}[dom])
            traffic_src = {"profile": ent.get("profile"), "commit": ent.get("commit"), "note": ent.get("source")}
        except Exception:
            traffic = None
    dom_units = hits if dom == "epa" else n                       # units one launch of that kernel processes (rank 0)
    dom_bytes = bytes_per_pair * dom_units
    achieved = dom_bytes / (kern[dom] * 1e-3) / 1e9 if kern[dom] > 0 else None
    step_ms = ms_total / args.steps
    cfg = static_config(desc, n, mode, world, bytes_per_pair, batch.verts.nbytes / 1e6)
    cfg["lanes_per_pair"] = args.lanes or "auto (hulls <= 16 vertices: 1 thread/pair, init kernel + compacting walk kernel; larger: 4 lanes/pair, 8 pairs per warp in lockstep)"
    cfg["host_numa_node"] = numa[0] if numa else None      # process + pinned buffers bound to the GPU's NUMA node
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "colliding_fraction": hits_all / pairs_all,
        "colliding_flags_per_s": hits_all * args.steps / (ms_total * 1e-3),
        "roofline": {"bound": "hbm", "kernel": {"gjk": "gjk stage (gjk_init_kernel + gjk_walk_kernel)", "epa": "epa stage (epa3_kernel + epa_kernel)"}[dom],
                     "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": (achieved / peak_gbs) if achieved else None, "peak_source": peak_src,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": dom_bytes,
                     "kernel_ms": kern, "kernel_share_of_step": kern[dom] / step_ms if step_ms else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": dt_copy / e2e_steps * 1e3, "steps": e2e_steps,
                "api": "gjkepa_batch_f64 (C ABI, pinned host buffers; everything the reference call takes crosses PCIe every step, "
                       "in chunks that overlap the kernels)",
                "h2d_gbs_per_gpu": h2d / (dt_copy / e2e_steps) / 1e9,
                "limiter": "PCIe: %.0f MB of world-space vertices per GPU per step; the kernels are %.0f %% of the step" % (
                    h2d / 1e6, 100.0 * step_ms / (dt_copy / e2e_steps * 1e3)),
                "resident_store": e2e_res},
        "pcie": pcie,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if e2e_solo:
        line["e2e"]["rank0_alone_value"] = e2e_solo
        line["e2e"]["efficiency_vs_rank0_alone"] = (e2e_value / world) / e2e_solo
        line["e2e"]["efficiency_note"] = ("per-GPU rate with all ranks copying at once over the rate of rank 0 alone on the same box: "
                                          "what the shared host memory / PCIe root costs (see pcie.*_all_ranks_at_once)")
    if gjk_epa:
        line["gjk_epa"] = gjk_epa
        line["contacts_per_s"] = gjk_epa["contacts_per_s"]      # GJK+EPA contacts (depth and normal produced) per second, all GPUs
    if one_process:
        line["one_process"] = one_process
    if not args.no_cpu:
        sample = min(n, 1_000_000)
        if numa:
            os.sched_setaffinity(0, numa[1])             # the CPU baseline gets every host core back
        rate, used, ns, _, dt = cpu_rate(batch, mode, sample, repeats=2)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
                                "sample": f"first {ns} pairs of the same workload, best of 2 ({dt:.2f} s); C restatement "
                                          "of GCLIB_GJKEPA.f90 (no Fortran compiler in the image), OpenMP over pairs"}
        if numa:
            _shard.bind_to_gpu_numa(local_rank)
    if not args.no_extra and args.workload == "cfg2" and world == 1:
        # the same kind of step with the shapes resident on the device and only poses crossing PCIe
        # (gjkepa_shapes_upload once, gjkepa_batch_posed_f64 per step): 2 bodies per pair, every body its own box
        pb = W.make_posed(n_shapes=2 * args.pairs, n_bodies=2 * args.pairs, n_pairs=args.pairs, sigma=1.0, kind="box", seed=777)
        eng.upload_shapes(pb.shape_verts, pb.shape_off)
        hp = [pin(x) for x in (pb.body_shape, pb.pose, pb.pair_a, pb.pair_b)]
        po = {"collision": pinned_empty((pb.n_pairs,), torch.uint8)}
        for _ in range(2):
            eng.batch_posed(*hp, mode="gjk", out=po)
        torch.cuda.synchronize()
        kp = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(kp):
            eng.batch_posed(*hp, mode="gjk", out=po)
        torch.cuda.synchronize()
        dtp = (time.perf_counter() - t0) / kp
        line.setdefault("extra", {})["e2e_resident_shapes"] = {
            "workload": "cfg2-shaped: 2M box bodies (shapes resident on the device), 1M body pairs, poses per step",
            "value": pb.n_pairs / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3,
            "h2d_bytes_per_step": int(sum(x.nbytes for x in hp)), "d2h_bytes_per_step": int(pb.n_pairs),
            "colliding_fraction": float(po["collision"].mean()),
            "api": "gjkepa_batch_posed_f64 (C ABI, pinned host buffers; pose applied on the device)"}
    if not args.no_extra and world == 1:
        # the same device-resident step with the opt-in shortcut E1 (GJKEPA_TUNE_EARLY_OUT: flags identical, exit paths of
        # separated pairs differ; the headline above is the parity build, shortcut off)
        eng.set_max_hull_vertices(int(np.diff(batch.hull_off).max()))
        eng.set_tuning(eng.TUNE_EARLY_OUT, 1)
        oe = None
        for _ in range(3):
            oe = eng.batch_device(*d_in, mode=mode, out=oe, want_status=False)
        torch.cuda.synchronize()
        ke = max(3, min(args.steps, 50))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(ke):
            oe = eng.batch_device(*d_in, mode=mode, out=oe, want_status=False)
        g1.record()
        torch.cuda.synchronize()
        eng.set_tuning(eng.TUNE_EARLY_OUT, 0)
        mse = g0.elapsed_time(g1) / ke
        line.setdefault("extra", {})["early_out_e1"] = {
            "workload": desc, "ms_per_step": mse, "value": batch.n_pairs / mse * 1e3, "unit": UNIT,
            "flags_identical_to_parity_build": bool(torch.equal(oe["collision"], out["collision"])),
            "note": "opt-in GJKEPA_TUNE_EARLY_OUT: the walk of a separated pair ends when a support value proves separation"}
    if not args.no_extra and args.workload == "cfg2" and world == 1:
        # fp32 mode: the same cfg2 step through gjkepa_batch_f32 (float hull store over PCIe, binary64 kernels on the device)
        v32 = pin(batch.verts.astype(np.float32))
        f_in = [v32, h_in[1], h_in[2], h_in[3]]
        f_out = {"collision": pinned_empty((batch.n_pairs,), torch.uint8)}
        flags64 = h_out["collision"].copy()              # what gjkepa_batch_f64 returned for the same pairs (e2e leg above)
        eng.set_max_hull_vertices(int(np.diff(batch.hull_off).max()))
        for _ in range(2):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        kf = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(kf):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        dtf = (time.perf_counter() - t0) / kf
        line.setdefault("extra", {})["e2e_f32_mode"] = {
            "workload": "cfg2 with the hull store as float32 at the boundary",
            "value": batch.n_pairs / dtf, "unit": UNIT, "ms_per_step": dtf * 1e3,
            "h2d_bytes_per_step": int(sum(x.nbytes for x in f_in)), "d2h_bytes_per_step": int(batch.n_pairs),
            "colliding_fraction": float(f_out["collision"].mean()),
            "flag_mismatch_rate_vs_f64": float(np.mean(f_out["collision"] != flags64)),
            "api": "gjkepa_batch_f32 (C ABI, pinned host buffers; device widens to binary64, same kernels)"}
        # ... and with binary32 ARITHMETIC as a filter in front (GJKEPA_TUNE_F32_FILTER): a float kernel proves pairs apart along
        # the axes and the line of centroids; only the rest runs the binary64 kernels.  Flags must be the same.
        flags_nofilter = f_out["collision"].copy()
        eng.set_tuning(eng.TUNE_F32_FILTER, 1)
        for _ in range(2):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(kf):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        dtff = (time.perf_counter() - t0) / kf
        tot_f, filt_f = eng.f32_filter_stats()
        eng.set_tuning(eng.TUNE_F32_FILTER, 0)
        line["extra"]["e2e_f32_mode"]["with_f32_filter"] = {
            "value": batch.n_pairs / dtff, "unit": UNIT, "ms_per_step": dtff * 1e3,
            "fraction_filtered": filt_f / max(tot_f, 1), "fraction_of_separated_pairs_filtered": filt_f / max(1, int((flags_nofilter == 0).sum())),
            "flag_mismatch_rate_vs_unfiltered": float(np.mean(f_out["collision"] != flags_nofilter)),
            "flag_mismatch_rate_vs_f64": float(np.mean(f_out["collision"] != flags64)),
            "note": "opt-in: binary32 separation test (3 axes + line of centroids, margin 1e-5 (1 + max|coordinate|)) before the binary64 "
                    "kernels; the exit path of a filtered pair reads EARLY_OUT"}
        # ... and with the GJK walk itself in binary32 ARITHMETIC (GJKEPA_TUNE_F32_ARITH): the float store is walked
        # as it arrived, no widening and no binary64 kernel in GJK mode.  Flags are not promised equal: the rate is reported.
        eng.set_tuning(eng.TUNE_F32_ARITH, 1)
        for _ in range(2):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(kf):
            eng.batch_host_f32(*f_in, mode="gjk", out=f_out)
        torch.cuda.synchronize()
        dtfa = (time.perf_counter() - t0) / kf
        eng.set_tuning(eng.TUNE_F32_ARITH, 0)
        fa = f_out["collision"]
        line["extra"]["e2e_f32_mode"]["with_f32_arithmetic"] = {
            "value": batch.n_pairs / dtfa, "unit": UNIT, "ms_per_step": dtfa * 1e3,
            "flag_mismatch_rate_vs_f64_kernels_on_the_same_floats": float(np.mean(fa != flags_nofilter)),
            "flags_extra": int(((fa > 0) & (flags_nofilter == 0)).sum()), "flags_missing": int(((fa == 0) & (flags_nofilter > 0)).sum()),
            "flag_mismatch_rate_vs_f64": float(np.mean(fa != flags64)),
            "note": "opt-in: the reference's flag path compiled with float operands from the source text of the binary64 kernels "
                    "(csrc/gjk_core_body.inc), thresholds scaled to binary32 (1e-8 -> 1e-5, 1e-12 -> 1e-7); differing pairs touch to "
                    "within that band (tests/test_gpu_f32_mode.py checks each with Qhull)"}
        eng.set_max_hull_vertices(0)
    if not args.no_extra and args.workload == "cfg2" and world == 1:
        # upstream of the path: candidate pairs built on the device (gjkepa_broadphase_posed_f64) for 1M posed bodies
        # instancing 1000 resident hull(16) shapes on a jittered lattice -- the cfg5 kind of scene
        rng = np.random.default_rng(99)
        nb = args.pairs
        ps = W.make_posed(n_shapes=1000, n_bodies=2, n_pairs=1, kind="hull", n_lo=16, n_hi=16, seed=5)
        eng.upload_shapes(ps.shape_verts, ps.shape_off)
        g = int(np.ceil(nb ** (1.0 / 3.0)))
        idx = np.arange(nb)
        site = np.stack([idx % g, (idx // g) % g, idx // (g * g)], axis=1).astype(np.float64)
        t = (site + rng.uniform(-0.3, 0.3, size=(nb, 3))) * 1.6
        pose = np.concatenate([W._quat_rot(rng, nb).reshape(nb, 9), t], axis=1)
        shp = rng.integers(0, 1000, size=nb).astype(np.int32)
        hb = [pin(x) for x in (shp, pose)]
        cap = 24 * nb
        ta = torch.empty((cap,), dtype=torch.int32, pin_memory=True); tb_ = torch.empty((cap,), dtype=torch.int32, pin_memory=True)
        keep += [ta, tb_]
        import ctypes as C
        tot = C.c_int64(0)
        vp = lambda a: C.c_void_p(a.ctypes.data)

        def bp_call():
            rc = eng._L.gjkepa_broadphase_posed_f64(eng._h, nb, vp(hb[0]), vp(hb[1]), cap, C.c_void_p(ta.data_ptr()),
                                                    C.c_void_p(tb_.data_ptr()), C.cast(C.byref(tot), C.c_void_p))
            eng._check(rc, "gjkepa_broadphase_posed_f64")
        for _ in range(2):
            bp_call()
        kb = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(kb):
            bp_call()
        dtb = (time.perf_counter() - t0) / kb
        line.setdefault("extra", {})["e2e_broadphase"] = {
            "workload": f"{nb} posed hull(16) bodies (1000 resident shapes) on a jittered lattice, pitch 1.6",
            "bodies_per_s": nb / dtb, "candidate_pairs": int(tot.value), "candidate_pairs_per_s": tot.value / dtb,
            "ms_per_step": dtb * 1e3, "h2d_bytes_per_step": int(sum(x.nbytes for x in hb)),
            "d2h_bytes_per_step": int(8 * min(tot.value, cap)),
            "api": "gjkepa_broadphase_posed_f64 (C ABI, pinned host buffers; pairs passing the reference's sphere test F:1185)"}
        # ... and the whole step in one call: poses in, contacts out (gjkepa_scene_posed_f64, GJK + EPA over the candidates)
        ccap = 6 * nb
        sa = torch.empty((ccap,), dtype=torch.int32, pin_memory=True); sb_ = torch.empty((ccap,), dtype=torch.int32, pin_memory=True)
        sd = torch.empty((ccap,), dtype=torch.float64, pin_memory=True); sn = torch.empty((ccap, 3), dtype=torch.float64, pin_memory=True)
        keep += [sa, sb_, sd, sn]
        ncand, ncon = C.c_int64(0), C.c_int64(0)

        def scene_call():
            rc = eng._L.gjkepa_scene_posed_f64(eng._h, 1, 1.0, nb, vp(hb[0]), vp(hb[1]), 1, ccap, C.c_void_p(sa.data_ptr()),
                                               C.c_void_p(sb_.data_ptr()), C.c_void_p(sd.data_ptr()), C.c_void_p(sn.data_ptr()),
                                               C.c_void_p(0), C.c_void_p(0), C.c_void_p(0),
                                               C.cast(C.byref(ncand), C.c_void_p), C.cast(C.byref(ncon), C.c_void_p))
            eng._check(rc, "gjkepa_scene_posed_f64")
        for _ in range(2):
            scene_call()
        t0 = time.perf_counter()
        for _ in range(kb):
            scene_call()
        dts = (time.perf_counter() - t0) / kb
        eng.set_tuning(eng.TUNE_EARLY_OUT, 1)            # the same step with shortcut E1 (most candidates are separated)
        nc0 = int(ncon.value)
        for _ in range(2):
            scene_call()
        t0 = time.perf_counter()
        for _ in range(kb):
            scene_call()
        dts1 = (time.perf_counter() - t0) / kb
        eng.set_tuning(eng.TUNE_EARLY_OUT, 0)
        nc1 = int(ncon.value)
        scene_call()                                     # back on the parity build for the numbers below
        line["extra"]["e2e_scene_step_early_out"] = {"ms_per_step": dts1 * 1e3, "contacts": nc1, "same_contact_count": nc1 == nc0}
        line["extra"]["e2e_scene_step"] = {
            "workload": f"{nb} posed hull(16) bodies: pose application + candidate pairs + GJK/EPA over them, contacts to the host",
            "bodies_per_s": nb / dts, "candidate_pairs": int(ncand.value), "contacts": int(ncon.value),
            "pairs_per_s": ncand.value / dts, "contacts_per_s": ncon.value / dts, "ms_per_step": dts * 1e3,
            "h2d_bytes_per_step": int(sum(x.nbytes for x in hb)), "d2h_bytes_per_step": int(40 * min(ncon.value, ccap)),
            "api": "gjkepa_scene_posed_f64 (C ABI, pinned host buffers)"}
    emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
