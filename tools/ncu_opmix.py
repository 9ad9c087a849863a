#!/usr/bin/env python3
"""Dynamic SASS opcode mix of one kernel in an .ncu-rep (warp instructions executed per opcode)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
for r in rows:
    if len(r) > 5 and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        d = dict(zip(hdr, r))
        src = d["Source"].strip()
        toks = src.split()
        if toks and toks[0].startswith("@"): toks = toks[1:]
        op = toks[0].split(".")[0] if toks else "?"
        if op in ("LDS", "STS", "LDG", "STG", "LDGSTS", "MUFU", "SHFL", "REDUX"): op = ".".join(toks[0].split(".")[:2])
        n = int(d["Instructions Executed"]); ops[op] += n; thr[op] += int(d["Thread Instructions Executed"]); samp[op] += int(d["# Samples"])
tot = sum(ops.values()) or 1; ts = sum(samp.values()) or 1
print(f"{kern}: {tot} warp instructions; static {sum(1 for r in rows if r and r[0].startswith('0x'))} SASS lines")
for op, n in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f"{op:14s} {n/tot*100:5.1f}%  samples {samp[op]/ts*100:5.1f}%  thr/inst {thr[op]/max(n,1):4.1f}")
