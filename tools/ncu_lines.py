#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line: warp instructions executed and
stall samples per line of our .cu file (top N)."""
import csv, collections, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
per = collections.defaultdict(lambda: [0, 0, ""])
hdr = None
for r in rows:
    if r and r[0] == "Line No" and len(r) > 8:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        n = int(r[ie] or 0); s = int(r[smp] or 0)
    except ValueError:
        continue
    line = int(r[0])
    per[line][0] += n; per[line][1] += s
    per[line][2] = r[1]
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
print("total warp-instr %d, samples %d" % (tot, tots))
for line, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d %6.2f%% inst %6.2f%% smp | %s" % (line, 100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(tots, 1), v[2].strip()[:110]))
