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


def run_reference(args, rank):
    if rank != 0:
        return
    batch, mode, desc = make_workload(args.workload, args.pairs, seed=12345)
    sample = min(batch.n_pairs, 250_000)
    for _ in range(args.warmup):
        cpu_rate(batch, mode, min(sample, 20_000))
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        rate, used, n, hits, dt = cpu_rate(batch, mode, sample)
        n_done += n
    total = time.perf_counter() - t0
    value = n_done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "pairs_per_step": sample, "mode": mode},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"first {sample} pairs of the workload per step; C restatement of "
                                   "GCLIB_GJKEPA.f90 (not gfortran: no Fortran compiler in the image), "
                                   "OpenMP parallel-for over pairs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from collision_detect_gjk_epa_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GJK/EPA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from collision_detect_gjk_epa_b200 import shard as _shard
    numa = _shard.bind_to_gpu_numa(local_rank)       # pinned buffers on the GPU's own NUMA node (undone before the CPU baseline)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch, mode, desc = make_workload(args.workload, args.pairs, seed=12345 + rank)
    eng = Engine(local_rank, lanes=args.lanes)
    eng.set_max_hull_vertices(int(np.diff(batch.hull_off).max()))      # the caller knows its hull sizes
    n = batch.n_pairs
    out_bytes = {"gjk": 1, "epa": 33, "full": 109}[mode]
    bytes_per_pair = batch.algorithmic_bytes_per_pair(out_bytes)

    # ---- device-resident inputs ----
    d_in = [torch.from_numpy(x).to(dev) for x in (batch.verts, batch.hull_off, batch.pair_a, batch.pair_b)]
    out = None

    def step_device():
        nonlocal out
        out = eng.batch_device(*d_in, mode=mode, out=out, want_status=False)

    for _ in range(max(3, args.warmup)):
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
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count()
    # per-kernel durations: one more (untimed) profiled step gives a per-launch sample of each kernel
    ksamples = []
    eng.set_profiling(True)
    for _ in range(5):
        step_device()
        ksamples.append(eng.kernel_ms())
    for k in kern:
        kern[k] = float(np.median([s[k] for s in ksamples]))
    eng.set_profiling(False)
    hits = int(out["collision"].sum().item())

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n), float(hits)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    pairs_all, hits_all = float(tot[0].item()), float(tot[1].item())
    value = pairs_all * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region) ----
    def pin(a):
        tt = torch.from_numpy(a).pin_memory()
        return tt, tt.numpy()
    keep = []
    h_in = []
    for x in (batch.verts, batch.hull_off, batch.pair_a, batch.pair_b):
        tt, npv = pin(x)
        keep.append(tt); h_in.append(npv)
    h_out = {}
    okeys = {"gjk": ["collision"], "epa": ["collision", "depth", "normal"],
             "full": ["collision", "depth", "normal", "colli_type", "nearest", "point"]}[mode]
    shapes = {"collision": ((n,), np.uint8), "depth": ((n,), np.float64), "normal": ((n, 3), np.float64),
              "colli_type": ((n,), np.int32), "nearest": ((n, 6), np.float64), "point": ((n, 3), np.float64)}
    for k in okeys:
        tt = torch.empty(shapes[k][0], dtype=getattr(torch, np.dtype(shapes[k][1]).name), pin_memory=True)
        keep.append(tt); h_out[k] = tt.numpy()
    h2d = int(sum(x.nbytes for x in h_in))
    d2h = int(sum(v.nbytes for v in h_out.values()))
    eng._L.gjkepa_use_own_stream(eng._h)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        eng.batch_host(*h_in, mode=mode, out=h_out)
    barrier()
    sampler.live.set()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.batch_host(*h_in, mode=mode, out=h_out)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()      # sampled through both timed regions (device-resident steps, then end-to-end steps)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = pairs_all * e2e_steps / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None
    tfile = ROOT / "profiles" / "ncu_traffic.json"          # dram bytes per launch from the committed ncu --set full captures
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(args.workload, {})
        except Exception:
            traffic = None
    peaks, peak_src = None, "fallback"
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text()); peak_src = "measured"
    peak_gbs = float(peaks["hbm_gbs"]) if peaks else 6650.0
    dom = "epa" if (mode != "gjk" and kern["epa"] > kern["gjk"]) else "gjk"
    dom_units = hits if dom == "epa" else n                       # units one launch of that kernel processes (rank 0)
    dom_bytes = bytes_per_pair * dom_units
    achieved = dom_bytes / (kern[dom] * 1e-3) / 1e9 if kern[dom] > 0 else None
    step_ms = ms_total / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "pairs_per_gpu_per_step": n, "mode": mode, "sharding": f"pairs by index x{world}",
                   "l2": "inputs larger than L2 (hull store %.0f MB per GPU)" % (batch.verts.nbytes / 1e6),
                   "algorithmic_bytes_per_pair": bytes_per_pair,
                   "colliding_fraction": hits_all / pairs_all, "lanes_per_pair": args.lanes or "auto (hulls <= 16 vertices: 1 thread/pair, init kernel + compacting walk kernel; larger: 4 lanes/pair, 8 pairs per warp in lockstep)"},
        "contacts_per_s": hits_all * args.steps / (ms_total * 1e-3),
        "roofline": {"bound": "hbm", "kernel": {"gjk": "gjk stage (gjk_init_kernel + gjk_walk_kernel)", "epa": "epa stage (epa3_kernel + epa_kernel)"}[dom],
                     "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": (achieved / peak_gbs) if achieved else None, "peak_source": peak_src,
                     "traffic": (traffic or {}).get({"gjk": "gjk_kernel", "epa": "epa_kernel"}[dom]),
                     "algorithmic_bytes_per_launch": dom_bytes,
                     "kernel_ms": kern, "kernel_share_of_step": kern[dom] / step_ms if step_ms else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "gjkepa_batch_f64 (C ABI, pinned host buffers)"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    line["config"]["host_numa_node"] = numa[0] if numa else None      # process + pinned buffers bound to the GPU's NUMA node
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
        # secondary headline: GJK+EPA contacts/s on cfg3 (1M pairs of 32-vertex hulls)
        b3, m3, d3 = make_workload("cfg3", args.pairs, seed=12345)
        t3 = [torch.from_numpy(x).to(dev) for x in (b3.verts, b3.hull_off, b3.pair_a, b3.pair_b)]
        o3 = None
        eng.set_max_hull_vertices(32)
        for _ in range(3):
            o3 = eng.batch_device(*t3, mode=m3, out=o3, want_status=False)
        torch.cuda.synchronize()
        k3 = max(3, min(args.steps, 20))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k3):
            o3 = eng.batch_device(*t3, mode=m3, out=o3, want_status=False)
        f1.record()
        torch.cuda.synchronize()
        ms3 = f0.elapsed_time(f1) / k3
        h3 = int(o3["collision"].sum().item())
        line["extra"] = {"workload": d3, "ms_per_step": ms3, "pairs_per_s": b3.n_pairs / ms3 * 1e3,
                         "contacts_per_s": h3 / ms3 * 1e3, "contacts": h3,
                         "algorithmic_gbs": b3.algorithmic_bytes_per_pair(33) * b3.n_pairs / (ms3 * 1e-3) / 1e9}
    if not args.no_extra and args.workload == "cfg2" and world == 1:
        # the same kind of step with the shapes resident on the device and only poses crossing PCIe
        # (gjkepa_shapes_upload once, gjkepa_batch_posed_f64 per step): 2 bodies per pair, every body its own box
        from collision_detect_gjk_epa_b200 import workloads as W
        pb = W.make_posed(n_shapes=2 * args.pairs, n_bodies=2 * args.pairs, n_pairs=args.pairs, sigma=1.0, kind="box", seed=777)
        eng.upload_shapes(pb.shape_verts, pb.shape_off)
        hp = []
        for x in (pb.body_shape, pb.pose, pb.pair_a, pb.pair_b):
            tt, npv = pin(x); keep.append(tt); hp.append(npv)
        tt = torch.empty((pb.n_pairs,), dtype=torch.uint8, pin_memory=True); keep.append(tt)
        po = {"collision": tt.numpy()}
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
        t32, v32 = pin(batch.verts.astype(np.float32)); keep.append(t32)
        f_in = [v32, h_in[1], h_in[2], h_in[3]]
        t32o = torch.empty((batch.n_pairs,), dtype=torch.uint8, pin_memory=True); keep.append(t32o)
        f_out = {"collision": t32o.numpy()}
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
        hb = []
        for x in (shp, pose):
            tt, npv = pin(x); keep.append(tt); hb.append(npv)
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
