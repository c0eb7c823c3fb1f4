"""TEST INFRASTRUCTURE: the non-headline methods on the synthetic headline-size scans (python tests/perf/big_methods.py bgkl,gp 65536)."""
import sys, time, json, numpy as np
sys.path.insert(0, "/root/repo")
import la3dm_b200
from la3dm_b200.synthetic import make_sequence
from oracle.ref import DEFAULT_PARAMS
FREE = {"bgk": 0.5, "bgkl": 0.3, "bgklv": 0.1, "gp": 0.1}
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
pts, org = make_sequence(4, npts, 50.0, 1)
for method in sys.argv[1].split(","):
    m = la3dm_b200.make_map(method, dict(DEFAULT_PARAMS[method]))
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        m.insert_pointcloud(pts[i], org[i], 0.1, FREE[method], -1.0 if method != "bgklv" else 30.0)
        ts.append(time.perf_counter() - t0)
        st = m.last_stats()
    print(json.dumps({"method": method, "points": npts, "ms": [round(1e3 * t, 2) for t in ts], "predict_ms": st["predict_ms"],
                      "n_train": st["n_train"], "tests": st["n_test_blocks"], "pairs": st["kernel_pairs"]}), flush=True)
    m.close()
