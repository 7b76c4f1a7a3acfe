#!/usr/bin/env python
"""Mnemonic counts per kernel of libcmcadi.so (static SASS): python tools/sass_digest.py > profiles/rNN_sass_digest.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "cmc_fluid_solver_b200" / "libcmcadi.so"
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
cols = ["UTMALDG", "SYNCS", "FENCE", "LDGSTS", "LDG", "LDG.256", "STG", "STG.256", "LDS", "STS", "BAR", "UCGABAR", "DFMA", "FFMA", "MUFU", "ATOMG"]
rows = {}
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        full = next(it)
        short = re.sub(r"\(.*", "", full).replace("void ", "").replace("cmc::", "").replace("(int)", "")
        cur = rows.setdefault(short, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base.startswith("UCGABAR"): cur["UCGABAR"] += 1
    elif base in ("ATOMG", "ATOM", "REDG", "RED"): cur["ATOMG"] += 1
    elif base == "LDG" and ".256" in op: cur["LDG.256"] += 1
    elif base == "STG" and ".256" in op: cur["STG.256"] += 1
    elif base in ("LDG", "STG", "LDS", "STS", "BAR", "DFMA", "FFMA", "MUFU", "ATOMG", "LDGSTS", "UTMALDG", "SYNCS", "FENCE", "UCGABAR"): cur[base] += 1
    elif base == "LD" or base == "ST":      # generic / strong accesses (volatile words of the one-pass slab exchange, peer stores)
        cur["LDG" if base == "LD" else "STG"] += 1
print("# SASS digest of cmc_fluid_solver_b200/libcmcadi.so (sm_100a), round 2 (final build)\n")
print("`tools/sass_digest.py`: `cuobjdump -sass cmc_fluid_solver_b200/libcmcadi.so`, instruction counts per kernel (static code, not executed counts).")
print("`UTMALDG` = `cp.async.bulk.tensor` (TMA tile loads), `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`, `UCGABAR` = cluster barrier,")
print("`LDG.256` / `STG.256` = 256-bit global accesses (`LDG.E.ENL2.256`), `FENCE` = `fence.proxy.async` / mbarrier-init fences.")
print("`k_tma_sweep<FT, DIR, chunks, lines, CTAs per tile, one-pass slab coupling>`.\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|" + "---|" * (len(cols) + 1))
for k in sorted(rows):
    if not any(rows[k][c] for c in cols):
        continue
    print(f"| `{k}` | " + " | ".join(str(rows[k][c]) for c in cols) + " |")
