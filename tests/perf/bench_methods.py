#!/usr/bin/env python
"""TEST INFRASTRUCTURE (times the checker under oracle/ next to the product; lives under tests/ for that reason).

Per-method timing on the reference's shipped sequences (BASELINE.json configs[1..3]): la3dm_b200 through the C ABI on
cuda:0 next to the reference's own sources (oracle/_ref, fast flavour, all host threads) on the same scans.

Not the headline benchmark (bench.py is); prints one JSON line per method for README / profiles.  Run on a GPU box:
    python tests/perf/bench_methods.py > gpurun_out/methods.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import la3dm_b200                      # noqa: E402
from oracle import ref                 # noqa: E402  (checker / CPU baseline only)

FREE_RES = {"bgk": 0.5, "bgkl": 0.3, "bgklv": 0.1, "gp": 0.1}     # config/methods/*.yaml
MAX_RANGE = 8.0                                                    # config/datasets/*.yaml

CASES = [
    # (method, fixture, scans, resolution, label)
    ("bgk", "scans_sim_unstructured.npz", 12, 0.1, "BGKOctoMap, sim_unstructured x12, res 0.1 (configs[1])"),
    ("bgkl", "scans_sim_structured.npz", 12, 0.1, "BGKLOctoMap, sim_structured x12, res 0.1"),
    ("bgklv", "scans_sim_structured.npz", 15, 0.05, "BGKLVOctoMap, sim_structured_long_term x15, res 0.05 (configs[2])"),
    ("gp", "scans_sim_unstructured.npz", 12, 0.1, "GPOctoMap, sim_unstructured x12, res 0.1 (configs[3])"),
]


def scans_for(method, fixture, n):
    z = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    pts, org = z["pts"], z["origins"]
    if method == "bgklv":      # data/sim_structured_long_term/*.pcd are copies of sim_structured_1.pcd
        return [(pts[0], org[0])] * n
    return [(pts[i], org[i]) for i in range(n)]


def run_gpu(method, params, scans, res):
    m = la3dm_b200.make_map(method, params)
    t = []
    visits = updates = 0
    for p, o in scans:
        t0 = time.perf_counter()
        m.insert_pointcloud(p, o, res, FREE_RES[method], MAX_RANGE)      # host cloud in, synchronous
        t.append(time.perf_counter() - t0)
        st = m.last_stats()
        visits += st["voxel_visits"]; updates += st["voxel_updates"]
    leaves = m.num_leaves()
    m.close()
    return t, leaves, visits, updates


def run_ref(method, params, scans, res):
    r = ref.RefMap(method, params, fast=True)
    t = []
    for p, o in scans:
        t0 = time.perf_counter()
        r.insert_pointcloud(p, o, res, FREE_RES[method], MAX_RANGE)
        t.append(time.perf_counter() - t0)
    n = len(r.leaves()["state"])
    threads = r.max_threads()
    r.close()
    return t, n, threads


def main():
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]
    no_ref = "--no-ref" in sys.argv[1:]
    for method, fixture, n, res, label in CASES:
        if only and method not in only:
            continue
        params = dict(ref.DEFAULT_PARAMS[method])
        params["resolution"] = res
        scans = scans_for(method, fixture, n)
        run_gpu(method, params, scans[:2], res)                         # warm-up: context, capacities, graph capture
        tg, leaves, visits, updates = run_gpu(method, params, scans, res)
        out = {"case": label, "method": method, "scans": n, "gpu_ms_total": 1e3 * sum(tg),
               "gpu_ms_per_scan_median": 1e3 * float(np.median(tg)), "gpu_ms_first_scan": 1e3 * tg[0],
               "leaves": int(leaves), "voxel_visits": int(visits), "voxel_updates": int(updates)}
        if not no_ref and ref.available(method, fast=True):
            tr, rleaves, threads = run_ref(method, params, scans, res)
            out.update({"ref_cpu_ms_total": 1e3 * sum(tr), "ref_cpu_ms_per_scan_median": 1e3 * float(np.median(tr)),
                        "ref_cpu_threads": int(threads), "ref_leaves": int(rleaves),
                        "speedup_total": sum(tr) / sum(tg)})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
