import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import la3dm_b200
from oracle import ref
from util import oracle_leaves_as_struct
P = dict(resolution=0.1, block_depth=5, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=0.2,
         prior_A=0.001, prior_B=0.001, original_size=True, min_W=0.001)
z = np.load("/root/repo/tests/golden/scans_sim_structured.npz")
pts, org = z["pts"][0], z["origins"][0]
m = la3dm_b200.BGKLVOctoMap(**P); r = ref.RefMap("bgklv", dict(P))
for s in range(2):
    m.insert_pointcloud(pts, org, 0.1, 0.1, 8.0); r.insert_pointcloud(pts, org, 0.1, 0.1, 8.0)
    g = m.leaves(); w = oracle_leaves_as_struct(r.leaves())
    print("scan", s, "stats", {k: v for k, v in m.last_stats().items() if k in ("n_train","n_test_blocks","voxel_visits","voxel_updates","kernel_pairs","n_blocks_total","new_blocks")})
    print(" leaves", len(g), len(w), "blocks", m.num_blocks(), r.num_blocks(), "unique keys", len(np.unique(g["block_key"])), len(np.unique(w["block_key"])))
    kg = set(np.unique(g["block_key"])); kw = set(np.unique(w["block_key"]))
    print(" blocks only ours", len(kg - kw), "only ref", len(kw - kg))
    for d in range(5):
        print("  depth", d, (g["depth"] == d).sum(), (w["depth"] == d).sum())
    for st in range(5):
        print("  state", st, (g["state"] == st).sum(), (w["state"] == st).sum())
    print("  classified", g["classified"].sum(), w["classified"].sum())
    if len(g) == len(w) and np.array_equal(g["block_key"], w["block_key"]) and np.array_equal(g["index"], w["index"]):
        rel = np.abs(g["prob"].astype(np.float64) - w["prob"]) / np.abs(w["prob"])
        print("  prob rel max", rel.max(), "n>1e-4", (rel > 1e-4).sum(), "a maxabs", np.abs(g["a"] - w["a"]).max(), "b maxabs", np.abs(g["b"] - w["b"]).max())
        i = int(np.argmax(rel)); print("  worst", g[i], w[i])
    else:
        # compare classified finest leaves as sets
        def cls(x): return set(zip(x["block_key"][x["classified"] == 1].tolist(), x["depth"][x["classified"] == 1].tolist(), x["index"][x["classified"] == 1].tolist()))
        cg, cw = cls(g), cls(w)
        print("  classified sets: ours", len(cg), "ref", len(cw), "only ours", len(cg - cw), "only ref", len(cw - cg))
        print("  sample only ref", list(cw - cg)[:5], "only ours", list(cg - cw)[:5])
