#!/usr/bin/env python3
"""Soak run on a GPU box: many seeds x workload shapes x kernel variants against the oracle.
Flags and GJK status words bit-exact, depths within 1e-9 relative.  Prints one line per case; exit code 1 on any mismatch."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from collision_detect_gjk_epa_b200 import Engine, workloads as W
from oracle import orc
from test_gpu_parity import configure

def main():
    seeds = range(int(os.environ.get("SEED0", "100")), int(os.environ.get("SEED0", "100")) + int(os.environ.get("SEEDS", "3")))
    eng = Engine(0)
    bad = 0
    t0 = time.time()
    for seed in seeds:
        cases = [("box8", W.make_pairs("box8", 200000, 1.0, seed=seed), 8),
                 ("hull12", W.make_pairs("hull", 120000, 0.7, seed=seed, n_vert=12), 12),
                 ("hull20", W.make_pairs("hull", 100000, 0.8, seed=seed, n_vert=20), 20),
                 ("hull32", W.make_pairs("hull", 100000, 0.85, seed=seed, n_vert=32), 32),
                 ("hull5", W.make_pairs("hull", 100000, 0.15, seed=seed, n_vert=5), 0),
                 ("mixed", W.make_mixed(60000, 0.4, seed=seed, n_lo=4, n_hi=64), 0),
                 ("touch", W.make_touching_mix(40000, seed=seed), 20),
                 ("store16", W.make_shared_store(30000, 150000, n_vert=16, seed=seed), 16)]
        for name, b, hint in cases:
            ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
            hit = (ref["collision"] > 0) & ((ref["status"] >> 24) == 0)
            for v in (0, "r1", "r32", "mono", "gen"):
                configure(eng, v)
                eng.set_max_hull_vertices(hint)
                got = eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
                f_ok = np.array_equal(got["collision"], ref["collision"])
                s_ok = np.array_equal(got["status"] & 0xffff, ref["status"] & 0xffff)
                fl_ok = bool(np.all((got["status"][hit] >> 24) == 0))
                dd = np.abs(got["depth"][hit] - ref["depth"][hit]) / np.maximum(ref["depth"][hit], 1.0)
                d_ok = bool(np.all(dd <= 1e-9))
                ok = f_ok and s_ok and fl_ok and d_ok
                bad += not ok
                print(f"seed {seed} {name:8s} variant {str(v):5s} flags {f_ok} status {s_ok} epa_flags {fl_ok} depth {d_ok} (max {dd.max() if dd.size else 0:.1e})"
                      + ("" if ok else "   <<<<<< MISMATCH"), flush=True)
    eng.close()
    print(f"soak done in {time.time() - t0:.0f} s, mismatching cases: {bad}")
    sys.exit(1 if bad else 0)

if __name__ == "__main__":
    main()
