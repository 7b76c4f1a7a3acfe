#!/usr/bin/env python
"""Top stall locations of one kernel in an .ncu-rep (SASS view): tools/ncu_hot.py rep kernel-regex [N] [launch-index]"""
import csv, subprocess, sys, io, re
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
# split per kernel
blocks = re.split(r'(?m)^"Kernel Name",', raw)[1:]
b = blocks[which]
lines = b.split("\n")
print("kernel:", lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stallcols}
print("stall totals:", ", ".join(f"{k[6:]} {v*100/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:topn]
for i in sorted(order):
    r = data[i]
    n = int(r[idx["# Samples"]])
    top = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stallcols), reverse=True)[:2]
    prev = data[i-1][idx["Source"]].strip()[:50] if i else ""
    print(f"{i:5d} {n*100/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
