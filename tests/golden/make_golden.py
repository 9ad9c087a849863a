#!/usr/bin/env python3
"""Generate tests/golden/gjkepa_golden.npz from the Python transliteration (ref_translit.py).

    python tests/golden/make_golden.py

The fixture holds the inputs (packed hull store + pair list) and, per pair, what the transliteration
of src/GCLIB_GJKEPA.f90 returns: collision flag, GJK exit path, walk iterations, support-call count
and -- with scipy/Qhull standing in for the missing GCLIB_QuickHull -- penetration depth and normal.
tests/test_golden.py checks the C oracle against it on the CPU, tests/test_gpu_parity.py the CUDA
path on the GPU.  Small (a few thousand pairs): the transliteration is pure-Python loops.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

import ref_translit as R  # noqa: E402
from collision_detect_gjk_epa_b200 import workloads as W  # noqa: E402

EXIT_CODE = {"sphere": 1, "init_cap": 2, "dup3": 3, "tri_origin": 4, "coplanar4": 5, "init_contain": 6,
             "loop_cap": 7, "loop_collinear": 8, "loop_coplanar": 9, "loop_contain": 10, "loop_stall": 11,
             "stop": 12}


def concat(batches):
    verts, offs, pa, pb, tag = [], [np.zeros(1, np.int64)], [], [], []
    hbase, vbase = 0, 0
    for t, b in enumerate(batches):
        verts.append(b.verts)
        offs.append(b.hull_off[1:] + vbase)
        pa.append(b.pair_a + hbase)
        pb.append(b.pair_b + hbase)
        tag.append(np.full(b.n_pairs, t, np.int32))
        hbase += b.n_hulls
        vbase += int(b.hull_off[-1])
    return W.PairBatch(np.concatenate(verts), np.concatenate(offs), np.concatenate(pa).astype(np.int32),
                       np.concatenate(pb).astype(np.int32), "golden"), np.concatenate(tag)


def main():
    cubes = [W.unit_cubes(o) for o in [(0.5, 0, 0), (0.5, 0.25, 0.125), (1.5, 0, 0), (1.0, 0, 0),
                                       (0.25, 0.5, 0.75), (0.9, 0.1, 0.2), (2.5, 0, 0), (0.0, 0.0, 0.999)]]
    names = ["cubes"] * len(cubes) + ["box8", "hull32", "mixed", "deep64", "hull5", "hull4", "hull16"]
    batches = cubes + [
        W.make_pairs("box8", 600, 1.0, seed=101),
        W.make_pairs("hull", 400, 0.85, seed=102, n_vert=32),
        W.make_mixed(300, 0.25, seed=103),
        W.make_pairs("hull", 60, 0.05, seed=104, n_vert=64),
        W.make_pairs("hull", 500, 0.15, seed=105, n_vert=5),
        W.make_pairs("hull", 400, 0.5, seed=106, n_vert=4),
        W.make_pairs("hull", 400, 0.85, seed=107, n_vert=16),
    ]
    b, tag = concat(batches)
    n = b.n_pairs
    col = np.zeros(n, np.uint8); ex = np.zeros(n, np.int32); gi = np.zeros(n, np.int32)
    ns = np.zeros(n, np.int32); ei = np.zeros(n, np.int32)
    depth = np.zeros(n); nrm = np.zeros((n, 3)); epa_ok = np.zeros(n, np.uint8)
    for i in range(n):
        r = R.GJKEPA(b.hull(int(b.pair_a[i])), b.hull(int(b.pair_b[i])))
        col[i] = r["collision"]; ex[i] = EXIT_CODE[r["exit"]]; gi[i] = r["gjk_iters"]; ns[i] = r["n_support"]
        if r["collision"] and r["note"] == "ok":
            epa_ok[i] = 1; depth[i] = r["depth"]; nrm[i] = r["normal"]; ei[i] = r["epa_iters"]
        if i % 500 == 0:
            print(i, "/", n, flush=True)
    np.savez_compressed(HERE / "gjkepa_golden.npz", verts=b.verts, hull_off=b.hull_off, pair_a=b.pair_a,
                        pair_b=b.pair_b, tag=tag, tag_names=np.array(names), collision=col, exit_path=ex,
                        gjk_iters=gi, n_support=ns, epa_ok=epa_ok, depth=depth, normal=nrm, epa_iters=ei)
    print("pairs", n, "colliding", int(col.sum()), "epa ok", int(epa_ok.sum()))


if __name__ == "__main__":
    main()
