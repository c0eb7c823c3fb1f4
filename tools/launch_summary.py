#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: total us, launches, share."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", r[ki].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    name = re.sub(r"^void ", "", name).replace("la3dm_b200::", "")
    name = re.sub(r"cub::CUB_\w+::", "cub::", name)[:48]
    d[name][0] += 1
    d[name][1] += v
tot = sum(v[1] for v in d.values())
print("%-50s %6s %10s %6s" % ("kernel", "n", "total_us", "share"))
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print("%-50s %6d %10.1f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("%-50s %6d %10.1f" % ("TOTAL", sum(v[0] for v in d.values()), tot))
