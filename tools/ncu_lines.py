#!/usr/bin/env python3
"""development helper: per-source-line instruction counts / stall samples from `ncu --page source --csv --print-source cuda`
usage: ncu -i rep --page source --csv --print-source cuda [--kernel-name ...] | tools/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
# the dump is a sequence of blocks: "File Path", "Function Name", header, lines...
blocks = []; cur = None
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = {"file": r[1], "rows": []}; blocks.append(cur); continue
    if r[0] == "Function Name": cur["func"] = r[1]; continue
    if r[0] == "Line No" or r[0] == "#": cur["hdr"] = r; continue
    if cur is not None: cur["rows"].append(r)
agg = {}
for b in blocks:
    h = b.get("hdr");
    if not h: continue
    ix = {n: i for i, n in enumerate(h)}
    key_fn = b["func"].split("(")[0][-40:]
    for r in b["rows"]:
        try:
            ln = r[0]; src = r[1]; inst = int(r[ix["Instructions Executed"]]); smp = int(r[ix["# Samples"]])
            tinst = int(r[ix["Thread Instructions Executed"]])
        except Exception: continue
        k = (key_fn, b["file"].split("/")[-1], ln)
        a = agg.setdefault(k, [0, 0, 0, src])
        a[0] += inst; a[1] += smp; a[2] += tinst
for fn in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ti = sum(v[0] for _, v in items); ts = sum(v[1] for _, v in items)
    print(f"== {fn}: inst {ti} samples {ts}")
    for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[1]:>16}:{k[2]:>4} inst {100*v[0]/max(ti,1):5.1f}% smp {100*v[1]/max(ts,1):5.1f}% thr/inst {v[2]/max(v[0],1):4.1f} | {v[3].strip()[:110]}")
