#!/usr/bin/env python3
"""Top source lines of one kernel in an .ncu-rep by executed instructions / stall samples."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr, r))
        try:
            agg.append((cur, int(r[0]), r[1][:100], int(d["Instructions Executed"]), int(d["# Samples"]),
                        int(d["Thread Instructions Executed"])))
        except Exception:
            pass
tot = sum(a[3] for a in agg) or 1; tots = sum(a[4] for a in agg) or 1
print(f"kernel {kern}: {tot} warp instructions, {tots} samples")
key = 4 if "--by-samples" in sys.argv else 3
agg.sort(key=lambda a: -a[key])
for a in agg[:top]:
    print(f"{a[0][:16]:16s} L{a[1]:4d} inst {a[3]/tot*100:5.1f}% samp {a[4]/tots*100:5.1f}% thr/inst {a[5]/max(a[3],1):4.1f} | {a[2]}")
