#!/usr/bin/env python3
"""Latency of the drop-in single-pair call (gjkepa_pair_f64: host arrays in, six outputs back) on a GPU box."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from collision_detect_gjk_epa_b200 import GJKEPA, workloads as W
for off in ((0.5, 0.25, 0.125), (1.5, 0.0, 0.0)):
    b = W.unit_cubes(off)
    p1, p2 = b.hull(0), b.hull(1)
    for _ in range(50): GJKEPA(1, 1.0, p1, p2)
    t0 = time.perf_counter()
    for _ in range(2000): r = GJKEPA(1, 1.0, p1, p2)
    dt = (time.perf_counter() - t0) / 2000
    print(f"offset {off}: collision={r[0]} depth={r[5]:.3f}  {dt * 1e6:.1f} us per CALL GJKEPA (full outputs)")
