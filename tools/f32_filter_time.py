#!/usr/bin/env python3
"""What the binary32 separation filter (GJKEPA_TUNE_F32_FILTER) buys on the device (developer tool): the fp32 entry on cfg2 /
cfg5-shaped batches with the filter off and on, one chunk, GJK stage time from the library's CUDA events; run it under
`ncu --metrics gpu__time_duration.sum` to see the filter kernel itself next to K1i / K1w."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from collision_detect_gjk_epa_b200 import Engine, workloads as W

eng = Engine(0)
eng.set_tuning(eng.TUNE_HOST_CHUNKS, 1)
for name, b in (("cfg2", W.make_pairs("box8", 1_000_000, 1.0)), ("cfg5", W.make_shared_store(100_000, 1_000_000, n_vert=16)),
                ("cfg3", W.make_pairs("hull", 500_000, 0.85, n_vert=32))):
    v32 = b.verts.astype(np.float32)
    eng.set_max_hull_vertices(int(np.diff(b.hull_off).max()))
    eng.set_profiling(True)
    res = {}
    for f in (0, 1):
        eng.set_tuning(eng.TUNE_F32_FILTER, f)
        ks = []
        for _ in range(5):
            out = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
            ks.append(eng.kernel_ms()["gjk"])
        res[f] = (float(np.median(ks[2:])), out["collision"].copy(), eng.f32_filter_stats())
    eng.set_profiling(False)
    tot, filt = res[1][2]
    sep = int((res[0][1] == 0).sum())
    print(f"{name}: GJK stage {res[0][0]:.3f} ms -> {res[1][0]:.3f} ms with the filter (filter kernel not included); filtered {filt} of {tot} pairs "
          f"= {filt / tot:.3f} ({filt / max(sep, 1):.3f} of the separated ones); flags equal: {np.array_equal(res[0][1], res[1][1])}", flush=True)
