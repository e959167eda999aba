#!/usr/bin/env python
"""Hottest SASS lines of one kernel from an `ncu --page source --csv --print-source sass` export (multi-kernel file).
usage: ncu_hot.py file.csv[.gz] <kernel substring> [nth match] [top N]"""
import csv, gzip, sys
path, pat = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
blocks, cur = [], None
for row in csv.reader(f):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
sel = [b for b in blocks if pat in b["name"]]
print("kernels matching:", len(sel), "of", len(blocks))
b = sel[nth]; h = {n: i for i, n in enumerate(b["hdr"])}
rows = b["rows"]
tot = sum(int(r[h["# Samples"]] or 0) for r in rows); ninst = sum(int(r[h["Instructions Executed"]] or 0) for r in rows)
print(b["name"][:120]); print("total samples", tot, "warp instructions", ninst, "sass lines", len(rows))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h["# Samples"]] or 0))[:top]
stalls = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
for i in sorted(order):
    r = rows[i]; s = int(r[h["# Samples"]] or 0)
    st = sorted(((int(r[h[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {100.0 * s / max(tot, 1):5.1f}% inst {int(r[h['Instructions Executed']] or 0):8d}  {r[h['Source']].strip()[:70]:70s} {st}")
# opcode histogram of executed warp instructions
import collections
hist = collections.Counter()
for r in rows:
    op = r[h["Source"]].strip().split()
    op = [x for x in op if not x.startswith("@")][:1]
    hist[(op[0].split(".")[0] if op else "?")] += int(r[h["Instructions Executed"]] or 0)
print("opcode mix:", ", ".join(f"{k} {100.0 * v / max(ninst, 1):.1f}%" for k, v in hist.most_common(14)))
