#!/usr/bin/env python3
"""One scene step (pose application, broad phase, GJK/EPA over the candidates, contact compaction) and one fp32-mode batch
at a size ncu can profile: developer tool for the Kp / Kb / Ks / Kc kernels (N bodies, default 1M)."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from collision_detect_gjk_epa_b200 import Engine, workloads as W

def main():
    nb = int(os.environ.get("N", "1000000")); reps = int(os.environ.get("REPS", "3"))
    rng = np.random.default_rng(99)
    ps = W.make_posed(n_shapes=1000, n_bodies=2, n_pairs=1, kind="hull", n_lo=16, n_hi=16, seed=5)
    g = int(np.ceil(nb ** (1.0 / 3.0))); idx = np.arange(nb)
    site = np.stack([idx % g, (idx // g) % g, idx // (g * g)], axis=1).astype(np.float64)
    t = (site + rng.uniform(-0.3, 0.3, size=(nb, 3))) * 1.6
    pose = np.concatenate([W._quat_rot(rng, nb).reshape(nb, 9), t], axis=1)
    shp = rng.integers(0, 1000, size=nb).astype(np.int32)
    eng = Engine(0)
    eng.set_tuning(eng.TUNE_EARLY_OUT, int(os.environ.get("E1", "0")))
    eng.upload_shapes(ps.shape_verts, ps.shape_off)
    for _ in range(reps):
        t0 = time.perf_counter()
        sc = eng.scene_posed(shp, pose, mode="epa", cap=6 * nb)
        dt = time.perf_counter() - t0
        print(f"[scene] {nb} bodies: {sc['n_candidates']} candidates, {sc['n_contacts']} contacts, {dt * 1e3:.2f} ms", flush=True)
    b = W.make_pairs("box8", nb, 1.0)
    v32 = b.verts.astype(np.float32)
    for _ in range(reps):
        t0 = time.perf_counter()
        out = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa")
        print(f"[f32] {nb} pairs: {(time.perf_counter() - t0) * 1e3:.2f} ms (pageable host buffers)", flush=True)

if __name__ == "__main__":
    main()
