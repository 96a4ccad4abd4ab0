#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass   (first kernel instance only)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
cur_file = None; hdr = None; col = None; seen_func = set(); skip = False
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == "": continue   # SASS row
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, [r[1].strip(), 0, 0])
    def num(x):
        try: return int(x)
        except ValueError: return 0
    a[1] += num(r[col["Instructions Executed"]])
    a[2] += num(r[col["# Samples"]])
ti = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("instructions", ti, "samples", ts)
topn = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f}:{l:4d} inst {100.0*a[1]/ti:5.1f}%  samp {100.0*a[2]/max(ts,1):5.1f}%  {a[0][:90]}")
