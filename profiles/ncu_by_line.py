#!/usr/bin/env python
"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by CUDA source line, using
nvdisasm line info of the cubin.  usage: ncu_by_line.py src.csv cubin kernel_substr [top]"""
import csv, re, subprocess, sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line = {}
infn, cur = False, None
for ln in dis:
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
base = None
tot = defaultdict(float)
keys = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared Excessive", "stall_short_sb", "stall_wait", "stall_barrier",
        "stall_mio", "stall_branch_resolving", "stall_no_inst", "stall_math", "stall_dispatch", "stall_long_sb", "stall_not_selected", "stall_selected"]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16)
    if base is None:
        base = a
    line = addr2line.get(a - base, (None, "?"))[0]
    for k in keys:
        try:
            v = float(r[col[k]] or 0)
        except ValueError:
            v = 0
        agg[line][k] += v
        tot[k] += v
print("TOTAL", {k: int(v) for k, v in tot.items()})
print(f"{'line':28s} " + " ".join(f"{k[-12:]:>12s}" for k in keys))
for line, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{str(line):28s} " + " ".join(f"{int(d[k]):12d}" for k in keys))
