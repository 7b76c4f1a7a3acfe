#!/usr/bin/env python
"""Stall samples of one kernel grouped into the code regions between barriers (SASS order):
tools/ncu_phases.py rep kernel-regex [launch-index]"""
import csv, subprocess, sys, io, re
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
b = re.split(r'(?m)^"Kernel Name",', raw)[1:][which]
lines = b.split("\n")
print("kernel:", lines[0][:100])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in data)
seg_start = 0
def flush(a, b):
    seg = data[a:b]
    n = sum(int(r[idx["# Samples"]]) for r in seg)
    if not seg: return
    agg = {h[6:]: sum(int(r[idx[h]] or 0) for r in seg) for h in stallcols}
    top = sorted(agg.items(), key=lambda x: -x[1])[:4]
    ops = {}
    for r in seg:
        op = r[idx["Source"]].strip().split()[0]
        if op.startswith("@"): op = r[idx["Source"]].strip().split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    keyops = {k: v for k, v in ops.items() if k in ("LDG", "STG", "LDGSTS", "LDS", "STS", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "DEPBAR", "LDGDEPBAR")}
    ex = sum(int(r[idx["Instructions Executed"]]) for r in seg)
    print(f"[{a:5d},{b:5d}) {n*100/tot:5.1f}%  exec {ex/1e6:7.1f}M  " + " ".join(f"{k}:{v*100/max(n,1):.0f}%" for k, v in top) + "   " + str(keyops))
for i, r in enumerate(data):
    s = r[idx["Source"]]
    if "BAR.SYNC" in s or "DEPBAR" in s:
        flush(seg_start, i + 1); seg_start = i + 1
flush(seg_start, len(data))
