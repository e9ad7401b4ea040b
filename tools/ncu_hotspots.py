#!/usr/bin/env python3
"""Aggregates the ncu source page (SASS view) of a report: share of stall samples by stall reason and by
opcode, plus the hottest instructions.  usage: ncu_hotspots.py report.ncu-rep [topN]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = raw.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
rs = Counter()
byop = Counter()
execs = Counter()
for r in rows:
    n = int(r["# Samples"] or 0)
    src = r["Source"].strip()
    t = src.split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    byop[op] += n
    execs[op] += int(r["Instructions Executed"] or 0)
    for k in reasons:
        rs[k] += int(r[k] or 0)
print("total samples", tot)
print("by reason:", {k: round(100.0 * v / tot, 1) for k, v in rs.most_common(10)})
print("samples by opcode (%):", {k: round(100.0 * v / tot, 1) for k, v in byop.most_common(14)})
te = sum(execs.values())
print("executed by opcode (%):", {k: round(100.0 * v / te, 1) for k, v in execs.most_common(16)})
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:top]:
    n = int(r["# Samples"] or 0)
    why = sorted(((int(r[k] or 0), k) for k in reasons), reverse=True)[:2]
    print(f"{100.0 * n / tot:5.2f}%  {r['Address']}  {r['Source'].strip()[:70]:70s} {why}")
