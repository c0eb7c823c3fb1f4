#!/usr/bin/env python
"""Per-source-line and per-opcode instruction counts of one kernel from an .ncu-rep (source page, SASS rows only)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, src, fname, ie, smp = None, {}, "", None, None
per = collections.defaultdict(lambda: [0, 0]); op = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1][:14]; continue
    if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r:
        ie = r.index("Instructions Executed"); smp = r.index("# Samples"); continue
    if ie is None or len(r) <= ie: continue
    if r[0].isdigit(): cur = (fname, int(r[0])); src[cur] = r[1]; continue
    if r[2] and r[2] != "-":
        try: n = int(r[ie] or 0); s = int(r[smp] or 0)
        except ValueError: continue
        per[cur][0] += n; per[cur][1] += s
        t = r[3].split(); o = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
        op[o.split(".")[0]] += n
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
print("total warp-instr %d samples %d" % (tot, tots))
for line, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %5d %6.2f%% inst %6.2f%% smp | %s" % (line[0], line[1], 100.0 * v[0] / tot, 100.0 * v[1] / max(tots, 1), src.get(line, "").strip()[:105]))
print(" ".join("%s:%.1f%%" % (o, 100.0 * n / tot) for o, n in op.most_common(24)))
