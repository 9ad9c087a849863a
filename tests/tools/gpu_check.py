#!/usr/bin/env python3
"""Developer check on a GPU box: parity of the CUDA path against the oracle on every workload
family + device-side timing of each mode.  Writes gpurun_out/gpu_check.json."""
import collections
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from collision_detect_gjk_epa_b200 import Engine, workloads as W  # noqa: E402
from collision_detect_gjk_epa_b200 import api  # noqa: E402
from oracle import orc  # noqa: E402


def parity(eng, b, mode, lanes_list=(4,), label=""):
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_FULL if mode == "full" else orc.MODE_EPA)
    res = {}
    for lanes in lanes_list:
        eng.set_lanes(lanes)
        got = eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode)
        flags_eq = bool(np.array_equal(got["collision"], ref["collision"]))
        gst, rst = got["status"], ref["status"]
        gjk_eq = bool(np.array_equal(gst & 0xffff, rst & 0xffff))
        hit = (ref["collision"] > 0) & (got["collision"] > 0)
        r = dict(flags_equal=flags_eq, gjk_status_equal=gjk_eq, n=int(b.n_pairs), hits=int(hit.sum()))
        if mode != "gjk" and hit.any():
            okst = ((gst >> 24) == 0) & ((rst >> 24) == 0)
            m = hit & okst
            r.update(gpu_flag_hist=dict(collections.Counter((gst[hit] >> 24).tolist())),
                     ref_flag_hist=dict(collections.Counter((rst[hit] >> 24).tolist())))
        if mode != "gjk" and hit.any() and m.any():
            dd = np.abs(got["depth"][m] - ref["depth"][m])
            tol = 1e-9 * np.maximum(ref["depth"][m], 1.0)
            cosd = np.abs(np.sum(got["normal"][m] * ref["normal"][m], axis=1))
            r.update(max_ddepth=float(dd.max()), depth_viol=int((dd > tol).sum()),
                     max_1mcos=float((1 - cosd).max()), normal_viol=int(((1 - cosd) > 1e-9).sum()),
                     gpu_flag_hist=dict(collections.Counter((gst[hit] >> 24).tolist())),
                     ref_flag_hist=dict(collections.Counter((rst[hit] >> 24).tolist())),
                     epa_iter_equal=float(np.mean(((gst >> 16) & 0xff)[m] == ((rst >> 16) & 0xff)[m])),
                     epa_iters_gpu_mean=float(((gst >> 16) & 0xff)[m].mean()),
                     epa_iters_ref_mean=float(((rst >> 16) & 0xff)[m].mean()))
            if mode == "full":
                r.update(type_equal=float(np.mean(got["colli_type"][m] == ref["colli_type"][m])),
                         nearest_maxdiff=float(np.abs(got["nearest"][m] - ref["nearest"][m]).max()),
                         nearest_equal=float(np.mean(np.all(got["nearest"][m] == ref["nearest"][m], axis=1))),
                         point_equal=float(np.mean(np.all(np.abs(got["point"][m] - ref["point"][m]) < 1e-9, axis=1))))
        res[f"lanes{lanes}"] = r
        print(f"[parity] {label or b.name} mode={mode} lanes={lanes}: {r}", flush=True)
    return res


def timing(eng, b, mode, lanes, iters=5):
    dev = torch.device("cuda:0")
    t = dict(verts=torch.from_numpy(b.verts).to(dev), hull_off=torch.from_numpy(b.hull_off).to(dev),
             pa=torch.from_numpy(b.pair_a).to(dev), pb=torch.from_numpy(b.pair_b).to(dev))
    eng.set_lanes(lanes)
    eng.set_max_hull_vertices(int(np.diff(b.hull_off).max()) if os.environ.get("HINT", "1") == "1" else 0)
    out = None
    times = []
    for i in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.batch_device(t["verts"], t["hull_off"], t["pa"], t["pb"], mode=mode, out=out, want_status=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[2:]))
    hits = int(out["collision"].sum().item())
    r = dict(ms=ms, pairs_per_s=b.n_pairs / ms * 1e3, contacts_per_s=hits / ms * 1e3, hits=hits, all_ms=times)
    print(f"[timing] {b.name} mode={mode} lanes={lanes}: {ms:.3f} ms  {r['pairs_per_s']:.3e} pairs/s  "
          f"{r['contacts_per_s']:.3e} contacts/s", flush=True)
    return r


def main():
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    print(torch.cuda.get_device_name(0), flush=True)
    eng = Engine(0)
    report = {"parity": {}, "timing": {}}
    lanes_all = (0, 1, 4)
    skip_parity = os.environ.get("SKIP_PARITY") == "1"
    # cfg1 cubes
    for off in [] if skip_parity else [(0.5, 0, 0), (0.5, 0.25, 0.125), (1.5, 0, 0), (1.0, 0, 0), (0.0, 0, 0), (0.5, 0.5, 0)]:
        b = W.unit_cubes(off)
        report["parity"][f"cubes{off}"] = parity(eng, b, "full", lanes_all, label=f"cubes{off}")
        col, typ, near, nrm, pt, dep = api.GJKEPA(1, 1.0, b.hull(0), b.hull(1))
        o = orc.pair(b.hull(0), b.hull(1))
        print(f"   GJKEPA{off}: gpu col={col} depth={dep} nrm={nrm} type={typ} pt={pt} | oracle col={o['collision']} "
              f"depth={o['depth']} nrm={o['normal']} type={o['colli_type']} pt={o['point']} st={hex(o['status'])}", flush=True)
    n = int(os.environ.get("CHECK_N", "200000"))
    sets = [
        ("cfg2", W.make_pairs("box8", n, 1.0), "gjk"),
        ("cfg2epa", W.make_pairs("box8", n // 4, 1.0, seed=3), "full"),
        ("cfg3", W.make_pairs("hull", n // 2, 0.85, n_vert=32), "full"),
        ("cfg4", W.make_mixed(n // 4, 0.25), "full"),
        ("cfg4deep", W.make_pairs("hull", n // 20, 0.05, n_vert=64), "full"),
        ("tiny5", W.make_pairs("hull", n // 2, 0.15, n_vert=5), "full"),
        ("tiny4", W.make_pairs("hull", n // 2, 0.5, n_vert=4), "full"),
        ("hull16", W.make_pairs("hull", n // 2, 0.85, n_vert=16), "epa"),
    ]
    for name, b, mode in ([] if skip_parity else sets):
        report["parity"][name] = parity(eng, b, mode, lanes_all, label=name)
    # timing at 1M pairs
    nt = int(os.environ.get("TIME_N", "1000000"))
    for name, b, modes in [
        ("cfg2", W.make_pairs("box8", nt, 1.0), ("gjk", "epa")),
        ("cfg3", W.make_pairs("hull", nt, 0.85, n_vert=32), ("gjk", "epa", "full")),
        ("cfg4", W.make_mixed(nt // 2, 0.25), ("epa",)),
        ("cfg4deep", W.make_pairs("hull", nt // 8, 0.05, n_vert=64), ("epa",)),
    ]:
        for mode in modes:
            for lanes in lanes_all:
                report["timing"][f"{name}_{mode}_L{lanes}"] = timing(eng, b, mode, lanes)
    (ROOT / "gpurun_out" / "gpu_check.json").write_text(json.dumps(report, indent=1, default=str))
    eng.close()


if __name__ == "__main__":
    main()
