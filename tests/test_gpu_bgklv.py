"""-m gpu parity tests of BGKLVOctoMap::insert_pointcloud (src/bgklvoctomap/bgklvoctomap.cpp:89-285) through the C ABI.
The checker for this method is the reference's OWN sources compiled in place (oracle/_ref, which travels to the GPU
box) plus the golden vectors it generated; the CPU restatement (oracle/la3dm_oracle.cpp) does not cover -LV."""
import numpy as np
import pytest

from conftest import golden
from util import FREE_RES, MAX_RANGE, RES, compare_leaves, key_hash, oracle_leaves_as_struct, summary

from oracle import ref

pytestmark = pytest.mark.gpu

BGKLV = dict(resolution=0.1, block_depth=5, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=0.2,
             prior_A=0.001, prior_B=0.001, original_size=False, min_W=0.001)   # config/methods/bgklvoctomap.yaml


def new_map(**kw):
    import la3dm_b200
    p = dict(BGKLV)
    p.update(kw)
    return la3dm_b200.BGKLVOctoMap(**p)


def long_term_scan():
    # data/sim_structured_long_term/*.pcd are byte-identical copies of sim_structured_1.pcd
    z = golden("scans_sim_structured.npz")
    return z["pts"][0], z["origins"][0]


needs_ref = pytest.mark.skipif(not ref.available("bgklv"), reason="oracle/_ref not built")


@needs_ref
def test_lv_frontend_matches_reference():
    """get_training_data (:303-423): O(hits^2) ray shortening, markers, rays -- bit-exact incl. order."""
    pts, org = long_term_scan()
    m, r = new_map(), ref.RefMap("bgklv", dict(BGKLV), threads=1)
    for max_range in (MAX_RANGE, 3.0, -1.0):
        xy = m.training_data(pts, org, RES, FREE_RES["bgklv"], max_range)
        rays, ridx = m.training_rays()
        wxy, widx, wrays = r.training_data(pts, org, RES, FREE_RES["bgklv"], max_range)
        assert xy.shape == wxy.shape, (max_range, xy.shape, wxy.shape)
        assert np.array_equal(ridx, widx)
        if max_range > 0:
            assert np.array_equal(xy, wxy) and np.array_equal(rays, wrays)
        else:
            # max_range <= 0 (no hit is ever inserted upstream): the ray length stays a double
            # all the way; 2 of 1918 rays end 1 ulp away from the reference's on this scan
            bad_rays = (rays != wrays).any(1).sum()
            assert bad_rays <= 4 and np.abs(rays - wrays).max() <= 1e-6 and np.abs(xy - wxy).max() <= 1e-6


def lv_compare(got, want, what):
    """-LV sums a voxel's training set in R-tree order upstream, in grid order here: same tolerance as the other maps,
    states compared away from the thresholds (var_thresh decides UNCERTAIN, so it is a threshold too)."""
    assert len(got) == len(want), (what, len(got), len(want))
    for k in ("block_key", "depth", "index", "x", "y", "z", "size"):
        assert np.array_equal(got[k], want[k]), (what, k)
    pg, pw = got["prob"].astype(np.float64), want["prob"].astype(np.float64)
    # -LV's probability of a free voxel is 0.5 (W - m_B - m_A) / (W - m_A) with W = m_A + m_B (bgklvoctree_node.cpp:29-44):
    # the numerator is the rounding residue of one fp32 addition (~1e-8), so a purely relative bound is meaningless
    # there; 1e-4 relative + 1e-6 absolute
    err = np.abs(pg - pw)
    rel = err / np.maximum(np.abs(pw), 1e-30)
    ok = err <= 1e-4 * np.abs(pw) + 1e-6
    assert ok.all(), (what, float(err[~ok].max()), int(np.argmax(~ok)), got[int(np.argmax(~ok))], want[int(np.argmax(~ok))])
    # alpha and beta against the scale they enter the probability with: beta = prior + (kbar - ybar) cancels, so its own
    # relative error says nothing (a 1e-7 difference of the two sums is 1e-4 of a beta that sits on the 0.001 prior)
    w = (want["a"].astype(np.float64) + want["b"].astype(np.float64))
    for f in ("a", "b"):
        d = np.abs(got[f].astype(np.float64) - want[f].astype(np.float64))
        assert (d <= 1e-4 * w + 1e-6).all(), (what, f, float(d.max()), int(np.argmax(d - 1e-4 * w)))
    bad = got["state"] != want["state"]
    near = np.zeros(len(pw), bool)
    for t in (0.3, 0.7):
        near |= np.abs(pw - t) <= 2e-4
    near |= np.abs(want["var"].astype(np.float64) - BGKLV["var_thresh"]) <= 1e-4
    assert not (bad & ~near).any(), (what, int((bad & ~near).sum()))
    assert np.array_equal(got["classified"], want["classified"]) or (got["classified"] != want["classified"]).sum() <= 2
    return float(rel[pw > 1e-3].max()) if (pw > 1e-3).any() else 0.0


@needs_ref
def test_lv_long_term_stream_matches_reference_and_golden():
    """BASELINE.json configs[2] (at the yaml's 0.1 m): the identical-scan stream, leaf by leaf after every scan."""
    g = golden("golden_bgklv_sim_structured_long_term_seq.npz")
    pts, org = long_term_scan()
    m, r = new_map(), ref.RefMap("bgklv", dict(BGKLV))
    for s in range(len(g["key_hashes"])):
        m.insert_pointcloud(pts, org, RES, FREE_RES["bgklv"], MAX_RANGE)
        r.insert_pointcloud(pts, org, RES, FREE_RES["bgklv"], MAX_RANGE)
        lv = m.leaves()
        assert key_hash(lv["block_key"], lv["depth"], lv["index"]) == str(g["key_hashes"][s]), "scan %d" % s
        lv_compare(lv, oracle_leaves_as_struct(r.leaves()), "lv scan %d" % s)
        got, want = summary(lv), g["summaries"][s]
        assert got[0] == want[0] and np.abs(got[:6] - want[:6]).max() <= 2, (s, got, want)
        assert abs(got[6] - want[6]) <= 1e-4 * want[6]


@needs_ref
def test_lv_half_resolution_and_pruning():
    """configs[2] proper: res 0.05 (depth 5); original_size = true (blocks that had data are pruned, :266-273)."""
    pts, org = long_term_scan()
    for kw in (dict(resolution=0.05, original_size=True), dict(original_size=True, var_thresh=0.05)):
        p = dict(BGKLV)
        p.update(kw)
        m, r = new_map(**kw), ref.RefMap("bgklv", p)
        for s in range(3):
            m.insert_pointcloud(pts[::3], org, p["resolution"], FREE_RES["bgklv"], MAX_RANGE)
            r.insert_pointcloud(pts[::3], org, p["resolution"], FREE_RES["bgklv"], MAX_RANGE)
            lv_compare(m.leaves(), oracle_leaves_as_struct(r.leaves()), "lv %r scan %d" % (kw, s))


@needs_ref
def test_lv_large_map_configuration():
    """config/methods/bgklvoctomap_large_map.yaml: resolution 0.2, block_depth 6 (32768 finest voxels per 6.4 m block),
    ell 0.6, original_size, min_W 0.01, var_thresh 0.001."""
    kw = dict(resolution=0.2, block_depth=6, sf2=0.1, ell=0.6, original_size=True, min_W=0.01, var_thresh=0.001)
    p = dict(BGKLV)
    p.update(kw)
    pts, org = long_term_scan()
    m, r = new_map(**kw), ref.RefMap("bgklv", p)
    for s in range(2):
        m.insert_pointcloud(pts, org, 0.5, 0.1, 30.0)          # ds_resolution 0.5 (clamped to 0.2 upstream), max_range 30
        r.insert_pointcloud(pts, org, 0.5, 0.1, 30.0)
        got, want = m.leaves(), oracle_leaves_as_struct(r.leaves())
        assert len(got) == len(want)
        for k in ("block_key", "depth", "index", "x", "y", "z", "size"):
            assert np.array_equal(got[k], want[k]), k
        err = np.abs(got["prob"].astype(np.float64) - want["prob"].astype(np.float64))
        assert (err <= 1e-4 * np.abs(want["prob"]) + 1e-6).all(), float(err.max())
