"""-m gpu tests of the rows next to the hot path: the batched point query
(BGKOctoMap::search, src/bgkoctomap/bgkoctomap.cpp:554-574 -> Block::search, bgkblock.cpp:132-156), block import
(inverse of the export behind begin_leaf()), and map save / load (checkpoint -> resume must continue bit for bit)."""
import numpy as np
import pytest

from util import FREE_RES, MAX_RANGE, RES

pytestmark = pytest.mark.gpu

BGK = dict(resolution=0.1, block_depth=3, sf2=1.0, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=100.0,
           prior_A=0.001, prior_B=0.001)   # config/methods/bgkoctomap.yaml
CMP = ("block_key", "depth", "index", "x", "y", "z", "size", "a", "b", "prob", "var", "state", "classified")


def new_map(method="bgk", **kw):
    import la3dm_b200
    from oracle.ref import DEFAULT_PARAMS
    p = dict(DEFAULT_PARAMS[method])
    p.update(kw)
    return la3dm_b200.make_map(method, p)


def build(scans, n=4, method="bgk", name="sim_structured", **kw):
    pts, org = scans[name]
    m = new_map(method, **kw)
    for i in range(n):
        m.insert_pointcloud(pts[i], org[i], kw.get("resolution", RES), FREE_RES[method], MAX_RANGE)
    return m


def leaf_holding(lv, q):
    """index into lv of the leaf whose half-open cube [c - s/2, c + s/2) holds q, or -1 (numpy reference)."""
    out = np.full(len(q), -1, np.int64)
    c = np.stack([lv["x"], lv["y"], lv["z"]], 1).astype(np.float64)
    h = lv["size"].astype(np.float64) * 0.5
    for i, p in enumerate(q.astype(np.float64)):
        inside = ((p >= c - h[:, None]) & (p < c + h[:, None])).all(1)
        j = np.flatnonzero(inside)
        if len(j):
            out[i] = j[0]
    return out


@pytest.mark.parametrize("depth", [3, 4, 2])
def test_search_returns_the_leaf_that_holds_the_point(scans, depth):
    m = build(scans, n=5, block_depth=depth)
    lv = m.leaves()
    rng = np.random.default_rng(7)
    # points inside known leaves (away from the faces: the cell index is a float division upstream), far-away points
    pick = rng.integers(0, len(lv), 4000)
    off = rng.uniform(-0.45, 0.45, (len(pick), 3)).astype(np.float32) * lv["size"][pick, None]
    q_in = np.stack([lv["x"][pick], lv["y"][pick], lv["z"][pick]], 1) + off
    q_far = rng.uniform(200.0, 300.0, (50, 3)).astype(np.float32)
    q = np.concatenate([q_in, q_far]).astype(np.float32)
    got = m.search(q)
    want = leaf_holding(lv, q)
    assert (want[:len(q_in)] >= 0).all() and (want[len(q_in):] < 0).all()
    found = got["depth"] >= 0
    assert np.array_equal(found, want >= 0)
    for k in CMP:
        assert np.array_equal(got[k][found], lv[k][want[found]]), k
    # a block that does not exist answers with the default node (upstream: `return OcTreeNode()`)
    miss = got[~found]
    assert (miss["state"] == 2).all() and (miss["classified"] == 0).all()
    assert np.allclose(miss["a"], BGK["prior_A"]) and np.allclose(miss["b"], BGK["prior_B"])
    keys = np.array([m.block_to_hash_key(*p) for p in q_far])
    assert np.array_equal(miss["block_key"], keys)


def test_search_finest_only_reports_pruned_nodes(scans):
    """finest_only reproduces what upstream's operator[] hands back: the finest-layer node, PRUNED inside a pruned leaf."""
    m = build(scans, n=12)
    lv = m.leaves()
    coarse = np.flatnonzero(lv["depth"] < 2)
    assert len(coarse) > 0                      # the 12-scan sequence prunes
    q = np.stack([lv["x"][coarse], lv["y"][coarse], lv["z"][coarse]], 1) + np.float32(0.01)
    fine = m.search(q, finest_only=True)
    assert (fine["depth"] == 2).all() and (fine["state"] == 3).all()          # PRUNED
    leaf = m.search(q)
    for k in CMP:
        assert np.array_equal(leaf[k], lv[k][coarse]), k


@pytest.mark.parametrize("method", ["bgk", "gp", "bgklv"])
def test_export_import_roundtrip_and_resume(scans, method, tmp_path):
    """checkpoint after 3 scans -> load into a fresh map -> 3 more scans == 6 scans without interruption, bit for bit"""
    name = "sim_unstructured" if method == "gp" else "sim_structured"
    pts, org = scans[name]
    kw = dict(block_depth=4) if method == "bgklv" else {}      # keep the -LV files small
    ins = lambda mp, i: mp.insert_pointcloud(pts[i], org[i], RES, FREE_RES[method], MAX_RANGE)
    a = new_map(method, **kw)
    for i in range(3):
        ins(a, i)
    f = str(tmp_path / "map.la3dm")
    a.save(f)
    b = new_map(method, **kw)
    b.load(f)
    assert b.num_blocks() == a.num_blocks()
    la, lb = a.leaves(), b.leaves()
    assert la.tobytes() == lb.tobytes()
    ka, na = a.blocks()
    kb, nb = b.blocks()
    assert np.array_equal(ka, kb) and na.tobytes() == nb.tobytes()
    # import_blocks is the same path without the file
    c = new_map(method, **kw)
    c.import_blocks(ka, na)
    assert c.leaves().tobytes() == la.tobytes()
    for i in range(3, 6):
        ins(a, i)
        ins(b, i)
    assert a.leaves().tobytes() == b.leaves().tobytes()
    assert a.last_stats()["voxel_visits"] == b.last_stats()["voxel_visits"]


def test_load_rejects_other_parameters_and_non_empty_maps(scans, tmp_path):
    import la3dm_b200
    a = build(scans, n=1)
    f = str(tmp_path / "map.la3dm")
    a.save(f)
    other = new_map("bgk", ell=0.3)
    with pytest.raises(la3dm_b200.La3dmError):
        other.load(f)
    with pytest.raises(la3dm_b200.La3dmError):
        a.load(f)                                # not empty
    with pytest.raises(la3dm_b200.La3dmError):
        new_map("bgk").load(str(tmp_path / "missing.la3dm"))
    g = tmp_path / "junk.la3dm"
    g.write_bytes(b"not a map file at all, definitely" * 8)
    with pytest.raises(la3dm_b200.La3dmError):
        new_map("bgk").load(str(g))
