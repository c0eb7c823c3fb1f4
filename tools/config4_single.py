#!/usr/bin/env python
"""BASELINE.json configs[4] on ONE GPU (262 144-point scans, 200 m, 0.05 m): per-scan device / predict time; meant to be
run under `ncu --metrics gpu__time_duration.sum` for the launch list of the non-predict part."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import la3dm_b200
from la3dm_b200.synthetic import make_sequence
k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = dict(resolution=0.05, block_depth=3, sf2=1.0, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=100.0,
         prior_A=0.001, prior_B=0.001)
pts, org = make_sequence(k, 262144, 200.0, seed=5)
d = [torch.from_numpy(pts[s]).cuda() for s in range(k)]
m = la3dm_b200.BGKOctoMap(device=0, **p)
m.reserve_blocks(150000000)
for s in range(k):
    m.insert_pointcloud(d[s], org[s], 0.05, 0.5, -1.0)
    st = m.last_stats()
    print("scan %d: device %.2f ms predict %.2f ms launches %d n_train %d tests %d replays %d" % (
        s, st["device_ms"], st["predict_ms"], st["kernel_launches"], st["n_train"], st["n_test_blocks"], st["replays"]), flush=True)
m.close()
