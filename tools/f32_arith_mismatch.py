#!/usr/bin/env python3
"""How often, and where, the binary32 walk (GJKEPA_TUNE_F32_ARITH) answers differently from the binary64 kernels on the same
float-rounded vertices (developer tool): BATCHES x 1M random pairs per workload, every differing pair measured with Qhull on its
Minkowski difference (binary64): positive = overlap depth, negative = minus a lower bound of the gap."""
import os
import sys
from pathlib import Path
import numpy as np
from scipy.spatial import ConvexHull
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from collision_detect_gjk_epa_b200 import Engine, workloads as W

batches = int(os.environ.get("BATCHES", "10"))
eng = Engine(0)
for name, make in (("box8 (cfg2)", lambda s: W.make_pairs("box8", 1_000_000, 1.0, seed=s)),
                   ("hull16 shared store (cfg5)", lambda s: W.make_shared_store(100_000, 1_000_000, n_vert=16, seed=s)),
                   ("hull32 (cfg3)", lambda s: W.make_pairs("hull", 500_000, 0.85, n_vert=32, seed=s))):
    n_tot = extra = missing = 0
    dist = []
    for k in range(batches):
        b = make(1000 + k)
        v32 = b.verts.astype(np.float32)
        eng.set_tuning(eng.TUNE_F32_ARITH, 0)
        f64 = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")["collision"].copy()
        eng.set_tuning(eng.TUNE_F32_ARITH, 1)
        f32 = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")["collision"].copy()
        eng.set_tuning(eng.TUNE_F32_ARITH, 0)
        d = np.nonzero(f32 != f64)[0]
        n_tot += b.n_pairs; extra += int(f32[d].sum()); missing += int((f32[d] == 0).sum())
        v = v32.astype(np.float64)
        for i in d:
            h = []
            for hn in (int(b.pair_a[i]), int(b.pair_b[i])):
                o0, o1 = int(b.hull_off[hn]), int(b.hull_off[hn + 1])
                h.append(v[3 * o0:3 * o1].reshape(3, o1 - o0).T)
            md = (h[0][:, None, :] - h[1][None, :, :]).reshape(-1, 3)
            dist.append(float((-ConvexHull(md).equations[:, 3]).min()))
    dist = np.array(dist)
    q = np.quantile(np.abs(dist), [0.5, 0.9, 1.0]).tolist() if len(dist) else []
    print(f"{name}: {n_tot} pairs, {extra + missing} flags differ ({(extra + missing) / n_tot:.2e}): {extra} extra, {missing} missing; "
          f"|origin to Minkowski hull| of the differing pairs: median / 90 % / max = {q}; outside the hull (separated, called colliding): "
          f"{int((dist < 0).sum())}, inside (overlapping, called separated): {int((dist > 0).sum())}", flush=True)
