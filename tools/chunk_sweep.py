#!/usr/bin/env python3
"""Host-buffer step time against the number of chunks per batch (developer tool): cfg2 with everything crossing PCIe,
cfg3 against a resident store with compact contact rows."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from collision_detect_gjk_epa_b200 import Engine, workloads as W

def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()

def timed(fn, k=10):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3

eng = Engine(0)
keep = []
b = W.make_pairs("box8", 1_000_000, 1.0)
h = []
for x in (b.verts, b.hull_off, b.pair_a, b.pair_b):
    t, v = pin(x); keep.append(t); h.append(v)
t, flags = pin(np.empty(b.n_pairs, np.uint8)); keep.append(t)
eng.set_max_hull_vertices(8)
for c in (1, 2, 4, 8, 16, 0):
    eng.set_tuning(eng.TUNE_HOST_CHUNKS, c)
    print(f"cfg2 copy-all chunks={c}: {timed(lambda: eng.batch_host(*h, mode='gjk', out={'collision': flags})):.3f} ms", flush=True)
eng.upload_store(h[0], h[1])
for c in (1, 2, 4, 8, 0):
    eng.set_tuning(eng.TUNE_HOST_CHUNKS, c)
    print(f"cfg2 resident chunks={c}: {timed(lambda: eng.batch_resident(h[2], h[3], mode='gjk', out={'collision': flags})):.3f} ms", flush=True)
b3 = W.make_pairs("hull", 1_000_000, 0.85, n_vert=32)
eng.set_max_hull_vertices(32)
eng.upload_store(b3.verts, b3.hull_off)
n = b3.n_pairs
out = {}
for k, (shape, dt) in {"collision": ((n,), np.uint8), "contact_index": ((n,), np.int32), "depth": ((n,), np.float64), "normal": ((n, 3), np.float64)}.items():
    t, v = pin(np.empty(shape, dt)); keep.append(t); out[k] = v
ta, pa = pin(b3.pair_a); tb, pb = pin(b3.pair_b); keep += [ta, tb]
for c in (1, 2, 4, 8, 0):
    eng.set_tuning(eng.TUNE_HOST_CHUNKS, c)
    print(f"cfg3 resident compact chunks={c}: {timed(lambda: eng.batch_resident(pa, pb, mode='epa', out=out, compact=True)):.3f} ms", flush=True)
h3 = []
for x in (b3.verts, b3.hull_off):
    t, v = pin(x); keep.append(t); h3.append(v)
for c in (1, 4, 8, 0):
    eng.set_tuning(eng.TUNE_HOST_CHUNKS, c)
    print(f"cfg3 copy-all compact chunks={c}: {timed(lambda: eng.batch_host(h3[0], h3[1], pa, pb, mode='epa', out=out, compact=True), 5):.3f} ms", flush=True)
