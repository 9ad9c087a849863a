#!/usr/bin/env python3
"""The binary32 arithmetic walk (GJKEPA_TUNE_F32_ARITH) on the fp32 entry (developer tool): flag mismatches against the binary64
kernels on the same float-rounded vertices and the end-to-end time of the call (host buffers, one chunk) with the mode off
and on.  Run it under `ncu --metrics gpu__time_duration.sum` to see K1f / K1fg next to K1i / K1w (REPS=1 then)."""
import os
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from collision_detect_gjk_epa_b200 import Engine, workloads as W

reps = int(os.environ.get("REPS", "7"))
eng = Engine(0)
eng.set_tuning(eng.TUNE_HOST_CHUNKS, int(os.environ.get("CHUNKS", "1")))
for name, b in (("cfg2", W.make_pairs("box8", 1_000_000, 1.0)), ("cfg5", W.make_shared_store(100_000, 1_000_000, n_vert=16)),
                ("cfg3", W.make_pairs("hull", 500_000, 0.85, n_vert=32)), ("cfg4", W.make_mixed(500_000, 0.25))):
    v32 = b.verts.astype(np.float32)
    eng.set_max_hull_vertices(0)
    for mode in ("gjk", "epa"):
        res = {}
        for arith in (0, 1):
            eng.set_tuning(eng.TUNE_F32_ARITH, arith)
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                out = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode=mode)
                ts.append((time.perf_counter() - t0) * 1e3)
            res[arith] = (float(np.median(ts[1:] or ts)), out)
        eng.set_tuning(eng.TUNE_F32_ARITH, 0)
        f64, f32 = res[0][1]["collision"], res[1][1]["collision"]
        fp = int(((f32 > 0) & (f64 == 0)).sum()); fn = int(((f32 == 0) & (f64 > 0)).sum())
        tot, decided = eng.f32_filter_stats()
        print(f"{name} {mode}: call {res[0][0]:.3f} ms -> {res[1][0]:.3f} ms with the binary32 walk; flags differing from the binary64 kernels: "
              f"{fp} extra + {fn} missing of {b.n_pairs} ({(fp + fn) / b.n_pairs:.1e}); colliding {float(f64.mean()):.3f}", flush=True)
