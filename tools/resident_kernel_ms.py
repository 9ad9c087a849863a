#!/usr/bin/env python3
"""Per-stage kernel time of one host batch against the resident store (envelope spheres computed at upload: the init
kernel reads them instead of computing them) next to the device-buffer call on the same pairs (developer tool)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from collision_detect_gjk_epa_b200 import Engine, workloads as W

eng = Engine(0)
for name, b, mode in (("cfg2", W.make_pairs("box8", 1_000_000, 1.0), "gjk"), ("hull16", W.make_pairs("hull", 1_000_000, 0.85, n_vert=16), "gjk"),
                      ("cfg3", W.make_pairs("hull", 1_000_000, 0.85, n_vert=32), "epa")):
    eng.set_max_hull_vertices(int(np.diff(b.hull_off).max()))
    eng.set_tuning(eng.TUNE_HOST_CHUNKS, 1)
    eng.upload_store(b.verts, b.hull_off)
    eng.set_profiling(True)
    ks = []
    for _ in range(6):
        eng.batch_resident(b.pair_a, b.pair_b, mode=mode, want_status=False)
        ks.append(eng.kernel_ms())
    res = {k: float(np.median([x[k] for x in ks[2:]])) for k in ks[0]}
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(x).to(dev) for x in (b.verts, b.hull_off, b.pair_a, b.pair_b)]
    ks, out = [], None
    for _ in range(6):
        out = eng.batch_device(*t, mode=mode, out=out, want_status=False)
        ks.append(eng.kernel_ms())
    devk = {k: float(np.median([x[k] for x in ks[2:]])) for k in ks[0]}
    eng.set_profiling(False)
    print(f"{name}: resident (spheres ready) {res}  |  device call (spheres fused) {devk}", flush=True)
