import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import la3dm_b200
from oracle import ref
from util import oracle_leaves_as_struct
P = dict(resolution=0.05, block_depth=5, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=0.2,
         prior_A=0.001, prior_B=0.001, original_size=True, min_W=0.001)
z = np.load("/root/repo/tests/golden/scans_sim_structured.npz")
pts, org = z["pts"][0][::3], z["origins"][0]
m = la3dm_b200.BGKLVOctoMap(**P); r = ref.RefMap("bgklv", dict(P))
m.insert_pointcloud(pts, org, 0.05, 0.1, 8.0); r.insert_pointcloud(pts, org, 0.05, 0.1, 8.0)
g = m.leaves(); w = oracle_leaves_as_struct(r.leaves())
def keyset(x): return {(int(k), int(d), int(i)): n for n, (k, d, i) in enumerate(zip(x["block_key"], x["depth"], x["index"]))}
kg, kw = keyset(g), keyset(w)
only_w = [k for k in kw if k not in kg]; only_g = [k for k in kg if k not in kw]
print("leaves", len(g), len(w), "only ref", len(only_w), "only ours", len(only_g))
from collections import Counter
print("only-ref depths", Counter(k[1] for k in only_w), "only-ours depths", Counter(k[1] for k in only_g))
# for coarse leaves only in ref: show our children
shown = 0
for k in only_w:
    if k[1] == 3 and shown < 4:
        kids = [(k[0], 4, k[2] * 8 + c) for c in range(8)]
        print("ref parent", k, "state", w["state"][kw[k]], "a,b", w["a"][kw[k]], w["b"][kw[k]])
        for kid in kids:
            if kid in kg:
                n = kg[kid]; print("   ours child", kid[2], "state", g["state"][n], "cls", g["classified"][n], "a,b", g["a"][n], g["b"][n], "prob", g["prob"][n], "var", g["var"][n])
        shown += 1
# common leaves stats
common = [k for k in kg if k in kw]
ig = np.array([kg[k] for k in common]); iw = np.array([kw[k] for k in common])
err = np.abs(g["prob"][ig].astype(np.float64) - w["prob"][iw])
print("common", len(common), "prob err max", err.max(), "state mismatches", int((g["state"][ig] != w["state"][iw]).sum()), "a err", np.abs(g["a"][ig]-w["a"][iw]).max(), "b err", np.abs(g["b"][ig]-w["b"][iw]).max())
bad = np.where(g["state"][ig] != w["state"][iw])[0][:5]
for b in bad: print("  mismatch", g[ig[b]], w[iw[b]])
