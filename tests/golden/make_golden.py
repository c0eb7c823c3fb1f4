#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run HERE (where /root/reference exists), after `make -C oracle ref`:

    python tests/golden/make_golden.py

Outputs are produced by oracle/_ref = the reference's own sources compiled in place (exact flavour, 1 thread) on the
reference's own shipped scans (data/sim_structured, data/sim_unstructured; sim_structured_long_term is 60 byte-identical
copies of sim_structured_1.pcd).  The scans themselves are stored too (float32 xyz + VIEWPOINT origin) because
/root/reference does not exist on the GPU box.  Parameters are config/methods/*.yaml + config/datasets/*.yaml and the
call is the static nodes' map.insert_pointcloud(cloud, origin, resolution, free_resolution, max_range)
(src/bgkoctomap/bgkoctomap_static_node.cpp:95).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref import RefMap, read_pcd, DEFAULT_PARAMS  # noqa: E402

REF = os.environ.get("LA3DM_REFERENCE", "/root/reference")
FREE_RES = {"bgk": 0.5, "bgkl": 0.3, "bgklv": 0.1, "gp": 0.1}   # config/methods/*.yaml free_resolution
MAX_RANGE = 8.0                                                  # config/datasets/*.yaml
RES = 0.1


def key_hash(lv):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(lv["block_key"]).tobytes())
    h.update(np.ascontiguousarray(lv["depth"].astype(np.int32)).tobytes())
    h.update(np.ascontiguousarray(lv["index"].astype(np.int32)).tobytes())
    return h.hexdigest()


def summary(lv):
    pv = lv["prob_var"].astype(np.float64)
    return np.array([len(lv["state"]), (lv["state"] == 0).sum(), (lv["state"] == 1).sum(), (lv["state"] == 2).sum(),
                     (lv["state"] >= 3).sum(), lv["classified"].sum(), pv[:, 0].sum(), pv[:, 1].sum()], np.float64)


def load_scans(name, n=12):
    pts, org = [], []
    for s in range(1, n + 1):
        p, o = read_pcd(os.path.join(REF, "data", name, "%s_%d.pcd" % (name, s)))
        pts.append(p)
        org.append(o)
    return np.stack(pts), np.stack(org)


def full_dump(lv):
    return dict(block_key=lv["block_key"], depth=lv["depth"].astype(np.uint8), index=lv["index"].astype(np.uint16),
                ab=lv["ab"], state=lv["state"], classified=lv["classified"])


def main():
    for name in ("sim_structured", "sim_unstructured"):
        pts, org = load_scans(name)
        np.savez_compressed(os.path.join(HERE, "scans_%s.npz" % name), pts=pts, origins=org)
        print(name, pts.shape)
    scans = {n: np.load(os.path.join(HERE, "scans_%s.npz" % n)) for n in ("sim_structured", "sim_unstructured")}

    # (1) single block-batch correctness case = BASELINE.json configs[0]: BGK, sim_structured scan 1, full dump
    m = RefMap("bgk", threads=1)
    pts, org = scans["sim_structured"]["pts"][0], scans["sim_structured"]["origins"][0]
    xy, _, _ = m.training_data(pts, org, RES, FREE_RES["bgk"], MAX_RANGE)
    m.insert_pointcloud(pts, org, RES, FREE_RES["bgk"], MAX_RANGE)
    lv = m.leaves()
    mn, mx = m.get_bbox()
    np.savez_compressed(os.path.join(HERE, "golden_bgk_sim_structured_scan1.npz"), train_xyzy=xy[:, [0, 1, 2, 6]],
                        bbox=np.stack([mn, mx]), summary=summary(lv), key_hash=key_hash(lv), **full_dump(lv))
    print("bgk scan1", summary(lv))

    # key / LUT known answers (src/bgkoctomap/bgkblock.cpp:7-32, 69-101) at depth 3, res 0.1
    rng = np.random.default_rng(7)
    q = np.concatenate([rng.uniform(-60, 60, (200, 3)), rng.integers(-150, 150, (56, 3)) * 0.2 + 0.2,
                        np.zeros((1, 3))]).astype(np.float32)
    keys = np.array([m.block_to_hash_key(*p) for p in q], np.int64)
    ext = np.stack([m.extended_block(k) for k in keys])
    cen = np.stack([m.hash_key_to_block(k) for k in keys])
    lut = np.stack([m.key_loc(d, i) for d in range(3) for i in range(8 ** d)])
    np.savez_compressed(os.path.join(HERE, "golden_keys_depth3_res0.1.npz"), xyz=q, keys=keys, extended=ext,
                        centers=cen, lut=lut)
    m.close()

    # (2) full sequences: per-scan summaries + final key hash + every-8th-leaf sample of the final map
    jobs = [("bgk", "sim_structured", 12), ("bgk", "sim_unstructured", 12), ("gp", "sim_unstructured", 12),
            ("bgkl", "sim_structured", 12), ("bgklv", "sim_structured_long_term", 3)]
    for method, ds, n in jobs:
        m = RefMap(method, threads=1)
        src = scans["sim_structured" if ds == "sim_structured_long_term" else ds]
        sums, hashes, ntrain = [], [], []
        for s in range(n):
            i = 0 if ds == "sim_structured_long_term" else s
            pts, org = src["pts"][i], src["origins"][i]
            ntrain.append(len(m.training_data(pts, org, RES, FREE_RES[method], MAX_RANGE)[0]))
            m.insert_pointcloud(pts, org, RES, FREE_RES[method], MAX_RANGE)
            lv = m.leaves()
            sums.append(summary(lv))
            hashes.append(key_hash(lv))
        sub = {k: v[::8] for k, v in full_dump(lv).items()}
        np.savez_compressed(os.path.join(HERE, "golden_%s_%s_seq.npz" % (method, ds)), summaries=np.stack(sums),
                            key_hashes=np.array(hashes), n_train=np.array(ntrain), n_leaves=len(lv["state"]),
                            params=np.array(list(DEFAULT_PARAMS[method].values()), np.float32), **sub)
        print(method, ds, sums[-1])
        m.close()


if __name__ == "__main__":
    main()
