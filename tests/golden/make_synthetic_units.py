#!/usr/bin/env python
"""Regenerates tests/golden/synthetic_units_seed1_64k_50m.json: per-scan unit counts (voxel visits / updates / kernel
pairs / training points / test blocks) of bench.py's default workload, counted by the CPU oracle
(oracle/la3dm_oracle.cpp) on the seeded synthetic sequence.  They are properties of the scan sequence; bench.py's
reference arm uses them as the numerator of its throughput and the GPU arm cross-checks its own counters against them.

    python tests/golden/make_synthetic_units.py [n_scans]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from la3dm_b200.synthetic import make_sequence  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    points, extent, seed = 65536, 50.0, 1
    pts, org = make_sequence(n, points, extent, seed)
    scans = bench.count_units_with_oracle(pts, org)
    with open(bench.UNITS_FIXTURE, "w") as f:
        json.dump({"points": points, "extent": extent, "seed": seed, "params": bench.BGK, "ds_resolution": bench.DS_RES,
                   "free_res": bench.FREE_RES, "max_range": bench.MAX_RANGE, "scans": scans}, f, indent=1)
    print("wrote", bench.UNITS_FIXTURE, len(scans))


if __name__ == "__main__":
    main()
