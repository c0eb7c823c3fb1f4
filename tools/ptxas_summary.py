#!/usr/bin/env python
"""Summarise la3dm_b200/csrc/build/*.ptxas.log: kernel, registers, spills, shared memory (our kernels only)."""
import glob, os, re, subprocess, sys
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "la3dm_b200", "csrc", "build")
for log in sorted(glob.glob(os.path.join(here, "*.ptxas.log"))):
    txt = open(log).read().split("Compiling entry function '")[1:]
    for blk in txt:
        name = blk.split("'")[0]
        if "cub" in name[:12]:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "")).split("::")[-1]
        regs = re.search(r"Used (\d+) registers", blk)
        spill = re.search(r"(\d+) bytes spill stores", blk)
        smem = re.search(r"(\d+) bytes smem", blk)
        print("%-22s %-28s regs=%-4s spill=%-5s smem=%s" % (os.path.basename(log)[:-10], short[:28], regs and regs.group(1),
                                                       spill and spill.group(1), smem and smem.group(1)))
