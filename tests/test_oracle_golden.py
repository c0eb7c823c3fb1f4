"""CPU: the oracle (oracle/la3dm_oracle.cpp, our restatement of the reference algorithm) against the committed golden
vectors, which were produced by the reference's own sources compiled in place (oracle/_ref, tests/golden/make_golden.py).
This is what pins the oracle; the -m gpu tests then compare the CUDA path with the oracle and the same vectors."""
import numpy as np
import pytest

from conftest import golden
from util import FREE_RES, MAX_RANGE, RES, key_hash, summary, oracle_leaves_as_struct, compare_leaves

from oracle import port, ref


def struct(o):
    return oracle_leaves_as_struct(o.leaves())


def test_key_and_lut_known_answers():
    """block_to_hash_key / hash_key_to_block / get_extended_block / init_key_loc_map (bgkblock.cpp:7-32, 69-101)."""
    g = golden("golden_keys_depth3_res0.1.npz")
    o = port.PortMap("bgk")
    keys = np.array([o.block_to_hash_key(*p) for p in g["xyz"]], np.int64)
    assert np.array_equal(keys, g["keys"])
    assert o.block_to_hash_key(0, 0, 0) == 576461302059761664
    ext = np.stack([o.extended_block(k) for k in g["keys"]])
    assert np.array_equal(ext, g["extended"])
    lut = np.stack([o.key_loc(d, i) for d in range(3) for i in range(8 ** d)])
    assert lut.shape == (73, 3) and np.array_equal(lut, g["lut"])
    assert set(np.unique(np.abs(lut[9:]))) == {np.float32(0.05), np.float32(0.15)}


def test_frontend_training_set_bit_exact(scans):
    g = golden("golden_bgk_sim_structured_scan1.npz")
    pts, org = scans["sim_structured"]
    o = port.PortMap("bgk")
    xy, _, _ = o.training_data(pts[0], org[0], RES, FREE_RES["bgk"], MAX_RANGE)
    assert np.array_equal(xy[:, [0, 1, 2, 6]], g["train_xyzy"])


def test_bgk_single_scan_full_dump(scans):
    """BASELINE.json configs[0]; known answers of SURVEY.md section 8c."""
    g = golden("golden_bgk_sim_structured_scan1.npz")
    pts, org = scans["sim_structured"]
    o = port.PortMap("bgk")
    o.insert_pointcloud(pts[0], org[0], RES, FREE_RES["bgk"], MAX_RANGE)
    lv = struct(o)
    assert len(lv) == 43100
    assert key_hash(lv["block_key"], lv["depth"], lv["index"]) == str(g["key_hash"])
    # same libm, but the reference sums a block's points in R-tree order, the port in training-set order
    ab, want = np.stack([lv["a"], lv["b"]], 1).astype(np.float64), g["ab"].astype(np.float64)
    assert np.abs(ab - want).max() <= 4e-6
    pg, pw = ab[:, 0] / ab.sum(1), want[:, 0] / want.sum(1)
    assert (np.abs(pg - pw) / pw).max() <= 1e-5
    assert np.array_equal(lv["state"], g["state"]) and np.array_equal(lv["classified"], g["classified"])
    s = summary(lv)
    assert list(s[:5]) == [43100, 5142, 3652, 34306, 0]        # SURVEY.md section 8c known answers
    assert np.array_equal(s[:6], g["summary"][:6])
    assert abs(s[6] - 20844.2408) < 1e-3
    st = o.last_stats()
    assert (st["n_train"], st["n_test_blocks"], st["voxel_visits"], st["pairs"]) == (5546, 764, 48896, 2484608)


@pytest.mark.parametrize("method,ds,n", [("bgk", "sim_structured", 12), ("bgk", "sim_unstructured", 12),
                                         ("bgkl", "sim_structured", 12), ("gp", "sim_unstructured", 12)])
def test_sequences_match_reference_summaries(scans, method, ds, n):
    g = golden("golden_%s_%s_seq.npz" % (method, ds))
    pts, org = scans[ds]
    o = port.PortMap(method)
    for s in range(n):
        o.insert_pointcloud(pts[s], org[s], RES, FREE_RES[method], MAX_RANGE)
        assert o.last_stats()["n_train"] == int(g["n_train"][s])
        lv = struct(o)
        assert key_hash(lv["block_key"], lv["depth"], lv["index"]) == str(g["key_hashes"][s]), (method, s)
        got, want = summary(lv), g["summaries"][s]
        if method == "gp":   # a probability within rounding of a threshold may classify differently
            assert got[0] == want[0] and np.abs(got[:6] - want[:6]).max() <= 2, (method, s, got, want)
        else:
            assert np.array_equal(got[:6], want[:6]), (method, s, got, want)
        # GP: fp32 Cholesky with cond ~1e4, plain-loop factorisation in both but different triangular-solve order
        assert abs(got[6] - want[6]) <= (1e-5 if method == "gp" else 1e-6) * abs(want[6])
    sub = lv[::8]
    assert np.array_equal(sub["block_key"], g["block_key"]) and np.array_equal(sub["index"], g["index"])
    if method == "gp":
        # var = sf2 - |L^-1 k|^2 cancels to ~1e-3 and enters as 1/var: two fp32 implementations of the same algorithm
        # agree only to ~1e-3 absolute in probability at the worst voxels ("GP conditioning")
        pw = 1.0 / (1.0 + np.exp(-0.1 * g["ab"][:, 0].astype(np.float64)))      # l / max_ivar = 100 / 1000
        err = np.abs(sub["prob"].astype(np.float64) - pw)
        assert np.percentile(err, 99) <= 2e-3 and err.max() <= 5e-2, (np.percentile(err, 99), err.max())
    else:
        np.testing.assert_allclose(np.stack([sub["a"], sub["b"]], 1), g["ab"], rtol=2e-4, atol=1e-5)


@pytest.mark.skipif(not ref.available("bgk"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("method", ["bgk", "bgkl", "gp"])
def test_port_against_compiled_reference_on_a_perturbed_scan(scans, method):
    """Not just the shipped scans: a jittered, re-centred copy run through both the compiled reference and the port."""
    pts, org = scans["sim_unstructured"]
    rng = np.random.default_rng(11)
    p = (pts[3] + rng.normal(scale=0.02, size=pts[3].shape)).astype(np.float32) + np.float32([3.3, -7.1, 0.4])
    o0 = org[3] + np.float32([3.3, -7.1, 0.4])
    r = ref.RefMap(method, threads=1)
    o = port.PortMap(method)
    for m in (r, o):
        m.insert_pointcloud(p, o0, RES, FREE_RES[method], MAX_RANGE)
        m.insert_pointcloud(p[::2], o0, RES, FREE_RES[method], MAX_RANGE)
    got, want = struct(o), oracle_leaves_as_struct(r.leaves())
    if method == "gp":      # see the conditioning note above: keys / leaf sets exact, probability close in absolute terms
        assert np.array_equal(got["block_key"], want["block_key"]) and np.array_equal(got["index"], want["index"])
        err = np.abs(got["prob"].astype(np.float64) - want["prob"])
        assert np.percentile(err, 99) <= 2e-3 and err.max() <= 5e-2
    else:                   # summation order inside a block differs (R-tree order vs training-set order)
        compare_leaves(got, want, prob_rtol=2e-5, what=method)


def test_edge_cases_empty_and_tiny_clouds():
    o = port.PortMap("bgk")
    o.insert_pointcloud(np.zeros((0, 3), np.float32), np.zeros(3, np.float32), RES, 0.5, MAX_RANGE)
    assert o.num_blocks() == 0 and o.last_stats()["n_train"] == 0
    # all points beyond max_range: filtered, nothing inserted (bgkoctomap.cpp:394-398, 230-232)
    far = np.float32([[100, 0, 0], [0, 100, 0]])
    o.insert_pointcloud(far, np.zeros(3, np.float32), RES, 0.5, MAX_RANGE)
    assert o.num_blocks() == 0
    # one point: one hit + beam samples
    o.insert_pointcloud(np.float32([[1.0, 0.2, 0.1]]), np.zeros(3, np.float32), RES, 0.5, MAX_RANGE)
    assert o.last_stats()["n_train"] == 5 and o.num_blocks() > 0   # hit + origin + d=0.5, 1.0 + tail l-0.5


def test_reference_search_semantics_at_depth_4():
    """Pins what la3dm_search restates (tests/test_gpu_query.py) on the reference's OWN search
    (src/bgkoctomap/bgkoctomap.cpp:554-574 -> Block::search, bgkblock.cpp:132-156) at block_depth 4, the one depth for
    which upstream's frozen Block::cell_num = 8 is right (bgkblock.cpp:105): the node returned for a point is the
    finest-layer node of the point's cell -- the leaf itself when the leaf is at the finest layer, a PRUNED node inside a
    pruned leaf -- and a default node where no block exists; `classified` is not copied out (bgkoctree_node.h:36-45)."""
    from oracle import ref
    if not ref.available("bgk"):
        pytest.skip("oracle/_ref not built")
    z = golden("scans_sim_structured.npz")
    p = dict(ref.DEFAULT_PARAMS["bgk"])
    p["block_depth"] = 4
    r = ref.RefMap("bgk", p, threads=1)
    for i in range(4):
        r.insert_pointcloud(z["pts"][i], z["origins"][i], 0.1, 0.5, 8.0)
    lv = r.leaves()
    rng = np.random.default_rng(3)
    fin = np.flatnonzero(lv["depth"] == 3)
    pick = rng.choice(fin, 20000, replace=False)
    off = rng.uniform(-0.45, 0.45, (len(pick), 3)).astype(np.float32) * lv["loc_size"][pick, 3:4]
    ab, st, cl = r.search(lv["loc_size"][pick, :3] + off)
    assert np.array_equal(ab, lv["ab"][pick]) and np.array_equal(st, lv["state"][pick])
    del cl      # the copy constructor leaves `classified` uninitialised (bgkoctree_node.h:36-45): indeterminate upstream
    coarse = np.flatnonzero(lv["depth"] < 3)
    assert len(coarse) > 0
    off = rng.uniform(-0.45, 0.45, (len(coarse), 3)).astype(np.float32) * lv["loc_size"][coarse, 3:4]
    ab, st, cl = r.search(lv["loc_size"][coarse, :3] + off)
    assert (st == 3).all()                                    # State::PRUNED
    ab, st, cl = r.search(np.array([[300.0, 300.0, 300.0]], np.float32))
    assert st[0] == 2 and np.allclose(ab[0], [p["prior_A"], p["prior_B"]])


@pytest.mark.skipif(not ref.available("bgk"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("method,depth", [("bgk", 4), ("bgkl", 5), ("gp", 4)])
def test_port_against_compiled_reference_at_the_large_map_depths(scans, method, depth):
    """The -m gpu tests of the `*_large_map.yaml` block depths (BGKL 5, GP 4; BGK 4) use the port as their checker:
    pin the port on the compiled reference at those depths first."""
    pts, org = scans["sim_structured" if method != "gp" else "sim_unstructured"]
    p = dict(ref.DEFAULT_PARAMS[method])
    p["block_depth"] = depth
    r, o = ref.RefMap(method, p, threads=1), port.PortMap(method, p)
    for s in range(2):
        for m in (r, o):
            m.insert_pointcloud(pts[s], org[s], RES, FREE_RES[method], MAX_RANGE)
    got, want = struct(o), oracle_leaves_as_struct(r.leaves())
    if method == "gp":
        assert np.array_equal(got["block_key"], want["block_key"]) and np.array_equal(got["index"], want["index"])
        err = np.abs(got["prob"].astype(np.float64) - want["prob"])
        assert np.percentile(err, 99) <= 2e-3 and err.max() <= 5e-2
    else:
        compare_leaves(got, want, prob_rtol=2e-5, what="%s depth %d" % (method, depth))


def test_sincos_restatement_matches_libm():
    """la3dm_b200/csrc/block_common.cuh:sincosf_libm restates glibc's sinf / cosf (double-precision polynomial pair of
    ARM optimized-routines, one rounding to float) so that the CUDA kernel values equal the compiled reference's bit
    for bit; this checks a numpy copy of the same arithmetic against the host libm on arguments of the form
    d * 2 * 3.1415926f, d in [0, 1) (bgkinference.h:115-116)."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    for f in (libm.sinf, libm.cosf):
        f.restype, f.argtypes = C.c_float, [C.c_float]
    H = float.fromhex
    hpi_inv, hpi = H("0x1.45F306DC9C883p+23"), H("0x1.921FB54442D18p0")
    c1, c2, c3, c4 = H("-0x1.ffffffd0c621cp-2"), H("0x1.55553e1068f19p-5"), H("-0x1.6c087e89a359dp-10"), H("0x1.99343027bf8c3p-16")
    s1, s2, s3 = H("-0x1.555545995a603p-3"), H("0x1.1107605230bc4p-7"), H("-0x1.994eb3774cf24p-13")
    rng = np.random.default_rng(5)
    d = np.concatenate([rng.random(60000), 1.0 - rng.random(20000) * 1e-3, rng.random(20000) * 1e-3]).astype(np.float32)
    y = (d * np.float32(2.0) * np.float32(3.1415926)).astype(np.float32)
    x0 = y.astype(np.float64)
    n = ((x0 * hpi_inv).astype(np.int32).astype(np.int64) + 0x800000) >> 24
    x = x0 - n * hpi
    x2 = x * x
    x = np.where(((n + 1) >> 1) & 1, -x, x)
    x3 = x * x2
    sp = (x + x3 * s1) + (x3 * x2) * (s2 + x2 * s3)
    x4 = x2 * x2
    cp = ((1.0 + x2 * c1) + x4 * c2) + (x4 * x2) * (c3 + x2 * c4)
    cp = np.where(n & 2, -cp, cp)
    sf, cf = sp.astype(np.float32), cp.astype(np.float32)
    got_s, got_c = np.where(n & 1, cf, sf), np.where(n & 1, sf, cf)
    want_s = np.array([libm.sinf(float(v)) for v in y], np.float32)
    want_c = np.array([libm.cosf(float(v)) for v in y], np.float32)
    assert np.array_equal(got_s, want_s) and np.array_equal(got_c, want_c)
