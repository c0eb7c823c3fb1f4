#!/usr/bin/env python
"""GPOctoMap on the shipped sim_unstructured sequence (configs[3]): per-scan wall / device / predict time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import la3dm_b200
z = np.load(os.path.join(ROOT, "tests", "golden", "scans_sim_unstructured.npz"))
kw = dict(resolution=0.1, block_depth=3, sf2=1.0, ell=1.0, noise=0.01, l=100.0, min_var=0.001, max_var=1000.0,
          max_known_var=0.02, free_thresh=0.3, occupied_thresh=0.7)
for rep in range(2):
    m = la3dm_b200.GPOctoMap(device=0, **kw)
    rows = []
    for i in range(12):
        t0 = time.perf_counter()
        m.insert_pointcloud(z["pts"][i], z["origins"][i], 0.1, 0.1, 8.0)
        w = (time.perf_counter() - t0) * 1e3
        st = m.last_stats()
        rows.append((w, st["device_ms"], st["predict_ms"], st["kernel_launches"], st["n_train"], st["n_data_blocks"], st["n_test_blocks"]))
    m.close()
print("wall_ms device_ms predict_ms launches n_train n_data n_test")
for r in rows:
    print("%.3f %.3f %.3f %d %d %d %d" % r)
