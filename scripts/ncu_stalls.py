#!/usr/bin/env python
"""Summarise an ncu --page source --csv dump: total stall samples by reason and the top SASS lines.
usage: ncu -i X.ncu-rep --page source --csv | python scripts/ncu_stalls.py [topN]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address": continue
    s = int(r[col["# Samples"]] or 0)
    per = {h: int(r[col[h]] or 0) for h in stall_cols}
    for h in stall_cols: tot[h] += per[h]
    lines.append((s, r[col["Source"]].strip(), per, int(r[col["Instructions Executed"]] or 0)))
total = sum(l[0] for l in lines)
print("total samples", total, " instructions executed", sum(l[3] for l in lines))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d} {100.0*v/max(total,1):5.1f}%")
topn = int(sys.argv[1]) if len(sys.argv) > 1 else 25
print("top lines:")
for idx, (s, src, per, ie) in sorted(enumerate(lines), key=lambda t: -t[1][0])[:topn]:
    top = sorted(per.items(), key=lambda kv: -kv[1])[:2]
    print(f"  #{idx:4d} {s:7d} {100.0*s/max(total,1):5.1f}%  {src[:70]:70s} {top[0][0]}={top[0][1]} {top[1][0]}={top[1][1]}")
