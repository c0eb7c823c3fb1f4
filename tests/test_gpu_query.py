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


def test_search_agrees_with_the_reference_at_depth_4(scans):
    """At block_depth 4 upstream's own search is right (Block::cell_num = 8, bgkblock.cpp:105), so the compiled
    reference is the checker: la3dm_search(finest_only) must name the same finest node, with the same state, as
    BGKOctoMap::search returns for the same points (tests/test_oracle_golden.py pins what upstream returns)."""
    from oracle import ref
    if not ref.available("bgk"):
        pytest.skip("oracle/_ref not built")
    from util import oracle_leaves_as_struct
    pts, org = scans["sim_structured"]
    p = dict(ref.DEFAULT_PARAMS["bgk"])
    p["block_depth"] = 4
    m, r = new_map("bgk", block_depth=4), ref.RefMap("bgk", p)
    for i in range(4):
        m.insert_pointcloud(pts[i], org[i], RES, FREE_RES["bgk"], MAX_RANGE)
        r.insert_pointcloud(pts[i], org[i], RES, FREE_RES["bgk"], MAX_RANGE)
    want = oracle_leaves_as_struct(r.leaves())
    rng = np.random.default_rng(11)
    fin = np.flatnonzero(want["depth"] == 3)
    pick = rng.choice(fin, 20000, replace=False)
    off = rng.uniform(-0.45, 0.45, (len(pick), 3)).astype(np.float32) * want["size"][pick, None]
    q = np.stack([want["x"][pick], want["y"][pick], want["z"][pick]], 1) + off
    ab, st, _ = r.search(q)
    assert np.array_equal(ab[:, 0], want["a"][pick]) and np.array_equal(st, want["state"][pick])
    got = m.search(q, finest_only=True)
    for k in ("block_key", "depth", "index"):
        assert np.array_equal(got[k], want[k][pick]), k
    # the parity bound is on the occupancy probability (1e-4 relative); alpha / beta themselves sit next to the 0.001
    # priors, where one ulp is already 1.2e-4 relative
    pw = ab[:, 0].astype(np.float64) / (ab[:, 0].astype(np.float64) + ab[:, 1])
    np.testing.assert_allclose(got["prob"], pw, rtol=1e-4)
    np.testing.assert_allclose(got["a"], ab[:, 0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["b"], ab[:, 1], rtol=1e-4, atol=1e-6)
    near = np.abs(want["prob"][pick] - 0.3) < 1e-4
    near |= np.abs(want["prob"][pick] - 0.7) < 1e-4
    assert np.array_equal(got["state"][~near], st[~near])
    # inside pruned leaves upstream hands back the PRUNED finest node; a missing block the default node
    coarse = np.flatnonzero(want["depth"] < 3)
    qc = np.stack([want["x"][coarse], want["y"][coarse], want["z"][coarse]], 1) + np.float32(0.01)
    _, stc, _ = r.search(qc)
    gc = m.search(qc, finest_only=True)
    assert (stc == 3).all() and (gc["state"] == 3).all() and (gc["depth"] == 3).all()
    far = np.array([[300.0, 300.0, 300.0]], np.float32)
    abf, stf, _ = r.search(far)
    gf = m.search(far, finest_only=True)
    assert gf["depth"][0] == -1 and gf["state"][0] == stf[0] == 2 and gf["a"][0] == abf[0, 0] and gf["b"][0] == abf[0, 1]


def test_raycast_agrees_with_the_reference_at_depth_4(scans):
    """BGKOctoMap::RayCaster (bgkoctomap.h:91-214) against the compiled reference at block_depth 4 (where upstream's
    `lim` and the frozen Block::cell_num agree): the same number of steps, and step by step the same point, block key,
    finest node, validity and node state; rays inside the map, leaving it, through unknown space, axis-aligned, with
    xy ties, and a ray whose start block does not exist."""
    from oracle import ref
    if not ref.available("bgk"):
        pytest.skip("oracle/_ref not built")
    pts, org = scans["sim_structured"]
    p = dict(ref.DEFAULT_PARAMS["bgk"])
    p["block_depth"] = 4
    m, r = new_map("bgk", block_depth=4), ref.RefMap("bgk", p)
    for i in range(3):
        m.insert_pointcloud(pts[i], org[i], RES, FREE_RES["bgk"], MAX_RANGE)
        r.insert_pointcloud(pts[i], org[i], RES, FREE_RES["bgk"], MAX_RANGE)
    rng = np.random.default_rng(3)
    o = org[0].astype(np.float32)
    ends = (o + rng.normal(size=(300, 3)) * np.float32([4, 4, 1])).astype(np.float32)
    starts = np.tile(o, (len(ends), 1)) + rng.normal(scale=0.3, size=(len(ends), 3)).astype(np.float32)
    extra_s = np.float32([[o[0], o[1], o[2]], [o[0], o[1], o[2]], [o[0] + 0.33, o[1] - 0.2, o[2]], [300, 300, 300],
                          [o[0], o[1], o[2]]])
    extra_e = np.float32([[o[0] + 3.05, o[1], o[2]], [o[0] + 2.0, o[1] + 2.0, o[2]], [o[0] - 2.5, o[1] + 2.5, o[2] + 0.4],
                          [301, 300, 300], [o[0] + 40.0, o[1] + 1.0, o[2] + 0.3]])
    starts, ends = np.concatenate([starts, extra_s]).astype(np.float32), np.concatenate([ends, extra_e]).astype(np.float32)
    steps, n = m.raycast(starts, ends, max_steps=600)
    total_valid = 0
    for k in range(len(starts)):
        w = r.raycast(starts[k], ends[k], max_steps=600)
        assert n[k] == len(w["valid"]), (k, n[k], len(w["valid"]))
        g = steps[k, :n[k]]
        assert np.array_equal(g["depth"] >= 0, w["valid"] == 1), k
        assert np.array_equal(g["block_key"], w["block_key"]), k
        assert np.array_equal((3 << 16) + g["index"], w["node_key"]), k           # node key = (depth << 16) + index
        assert np.array_equal(np.stack([g["x"], g["y"], g["z"]], 1), w["p"]), k
        v = w["valid"] == 1
        assert np.array_equal(g["state"][v], w["state"][v]), k
        np.testing.assert_allclose(g["a"][v], w["ab"][v, 0], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(g["b"][v], w["ab"][v, 1], rtol=1e-4, atol=1e-6)
        total_valid += int(v.sum())
    assert total_valid > 5000 and n[-2] == 0 and (steps[-1, :n[-1]]["depth"] == -1).any()


def test_raycast_depth_3_is_consistent_with_search(scans):
    """block_depth 3 (the reference configuration; upstream's RayCaster mixes 2^(depth-1) with cell_num = 8 there):
    every valid step is the finest node la3dm_search(finest_only) returns for the step's own point."""
    m = build(scans, n=3)
    o = scans["sim_structured"][1][0]
    rng = np.random.default_rng(5)
    ends = (o + rng.normal(size=(64, 3)) * np.float32([3, 3, 1])).astype(np.float32)
    steps, n = m.raycast(np.tile(o, (64, 1)), ends, max_steps=400)
    assert (n > 0).all()
    for k in range(64):
        g = steps[k, :n[k]]
        g = g[g["depth"] >= 0]
        s = m.search(np.stack([g["x"], g["y"], g["z"]], 1), finest_only=True)
        assert s.tobytes() == g.tobytes()


def test_incremental_export_mirrors_the_map(scans):
    """la3dm_export_touched: the server loop's mirror (bgkoctomap_server.cpp:94-144 keeps the OCCUPIED / FREE leaves of
    the whole map) maintained from the blocks each scan touched; after every scan it must equal the filtered full
    export.  Also: nothing to report without a scan in between, and a loaded map reports everything."""
    import la3dm_b200
    from test_gpu_bgk import new_map
    pts, org = scans["sim_unstructured"]
    m = new_map()
    mask = (1 << 0) | (1 << 1)                       # LA3DM_FREE | LA3DM_OCCUPIED (include/la3dm_b200.h)
    mirror = {}
    for s in range(6):
        m.insert_pointcloud(pts[s], org[s], RES, FREE_RES["bgk"], MAX_RANGE)
        st = m.last_stats()
        keys, lv = m.touched_leaves(mask)
        assert len(keys) == st["n_test_blocks"], (len(keys), st["n_test_blocks"])
        assert np.all(np.diff(keys) > 0)
        assert np.isin(lv["block_key"], keys).all() and np.isin(lv["state"], (0, 1)).all()
        for k in keys:
            mirror.pop(int(k), None)
        for k in np.unique(lv["block_key"]):
            mirror[int(k)] = lv[lv["block_key"] == k]
        full = m.leaves()
        want = full[np.isin(full["state"], (0, 1))]
        got = np.concatenate([mirror[k] for k in sorted(mirror)]) if mirror else want[:0]
        assert got.tobytes() == want.tobytes(), s
        k2, l2 = m.touched_leaves(mask)
        assert len(k2) == 0 and len(l2) == 0
    # unfiltered, not clearing: twice the same
    m.insert_pointcloud(pts[6], org[6], RES, FREE_RES["bgk"], MAX_RANGE)
    ka, la = m.touched_leaves(0xFF, clear=False)
    kb, lb = m.touched_leaves(0xFF, clear=True)
    assert np.array_equal(ka, kb) and la.tobytes() == lb.tobytes() and len(la) > 0
    full = m.leaves()
    assert la.tobytes() == full[np.isin(full["block_key"], ka)].tobytes()
    # a map filled by import reports every block once
    keys, nodes = m.blocks()
    m2 = new_map()
    m2.import_blocks(keys, nodes)
    kc, lc = m2.touched_leaves(0xFF)
    assert np.array_equal(kc, keys) and lc.tobytes() == full.tobytes()
