#!/usr/bin/env python3
"""Quick device-side timing of the GJK kernels on one workload (developer tool)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import torch
from collision_detect_gjk_epa_b200 import Engine, workloads as W
from oracle import orc

def main():
    wl = os.environ.get("WL", "cfg2"); n = int(os.environ.get("N", "1000000")); mode = os.environ.get("MODE", "gjk")
    lanes_list = [int(x) for x in os.environ.get("LANES", "0,1,4").split(",")]
    if wl == "cfg2": b = W.make_pairs("box8", n, 1.0)
    elif wl == "cfg3": b = W.make_pairs("hull", n, 0.85, n_vert=32)
    elif wl == "hull16": b = W.make_pairs("hull", n, 0.85, n_vert=16)
    elif wl == "cfg5": b = W.make_shared_store(n // 10, n, n_vert=16)
    elif wl == "cfg5big": b = W.make_shared_store(n // 2, n, n_vert=16)         # store larger than L2 (n = 4M: 768 MB)
    elif wl == "cfg5bigsorted":                                                   # the same pairs in broad-phase order (grouped by first body)
        b = W.make_shared_store(n // 2, n, n_vert=16)
        o = np.argsort(b.pair_a, kind="stable"); b.pair_a = np.ascontiguousarray(b.pair_a[o]); b.pair_b = np.ascontiguousarray(b.pair_b[o])
    elif wl == "cfg4": b = W.make_mixed(n, 0.25)
    elif wl == "cfg4deep": b = W.make_pairs("hull", n, 0.05, n_vert=64)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(x).to(dev) for x in (b.verts, b.hull_off, b.pair_a, b.pair_b)]
    eng = Engine(0)
    eng.set_max_hull_vertices(int(np.diff(b.hull_off).max()))
    eng.set_profiling(True)
    eng.set_tuning(eng.TUNE_SINGLE_KERNEL, int(os.environ.get('SINGLE', '0')))
    eng.set_tuning(eng.TUNE_WALK_REFILL, int(os.environ.get('REFILL', '8')))
    eng.set_tuning(eng.TUNE_EPA_GENERAL, int(os.environ.get('EPAGEN', '0')))
    eng.set_tuning(eng.TUNE_TILE_MAX_VERTICES, int(os.environ.get('TILEMAX', '16')))
    eng.set_tuning(eng.TUNE_EARLY_OUT, int(os.environ.get('E1', '0')))
    chk = int(os.environ.get("CHECK", "20000"))
    ref = orc.batch(b.verts, b.hull_off, b.pair_a[:chk], b.pair_b[:chk], mode=orc.MODE_GJK) if chk else None
    e1 = int(os.environ.get('E1', '0'))
    for lanes in lanes_list:
        eng.set_lanes(lanes)
        out = None; ks = []; tot = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.batch_device(*t, mode=mode, out=out)
            e1.record(); torch.cuda.synchronize()
            tot.append(e0.elapsed_time(e1)); ks.append(eng.kernel_ms())
        k = {key: float(np.median([x[key] for x in ks[3:]])) for key in ks[0]}
        ok = ""
        if ref is not None:
            got = out["collision"][:chk].cpu().numpy(); st = out["status"][:chk].cpu().numpy().view(np.uint32)
            ok = f" flags_ok={np.array_equal(got, ref['collision'])} status_ok={np.array_equal(st & 0xffff, ref['status'] & 0xffff)}"
        stall = out["status"].cpu().numpy().view(np.uint32)
        ok += f" cold_math={np.mean((stall >> 30) & 1):.4f}"
        print(f"[quick] {wl} mode={mode} lanes={lanes}: total {np.median(tot[3:]):.3f} ms  K0 {k['hull_prep']:.3f}  K1 {k['gjk']:.3f}  K3 {k['epa']:.3f}  "
              f"-> {b.n_pairs / np.median(tot[3:]) * 1e3:.3e} pairs/s{ok}", flush=True)

if __name__ == "__main__":
    main()
