#!/usr/bin/env python
"""Split an ncu SASS source page into barrier-delimited phases: samples / instructions between BAR.SYNCs.
usage: ncu_by_phase.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
phase, cur = [], dict(samples=0, inst=0, first=None, ops={})
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]]; s = int(r[col["# Samples"]] or 0); n = int(r[col["Instructions Executed"]] or 0)
    if cur["first"] is None: cur["first"] = r[col["Address"]]
    cur["samples"] += s; cur["inst"] += n; tot += s
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + n
    if "BAR.SYNC" in src:
        phase.append(cur); cur = dict(samples=0, inst=0, first=None, ops={})
phase.append(cur)
print("total samples", tot)
for i, p in enumerate(phase):
    if p["samples"] == 0 and p["inst"] == 0: continue
    top = sorted(p["ops"].items(), key=lambda kv: -kv[1])[:5]
    print(f"phase {i:3d} @{p['first']}: samples {p['samples']:5d} ({100*p['samples']/tot:5.1f}%)  warp-inst {p['inst']:9d}  " +
          " ".join(f"{k}:{v}" for k, v in top))
