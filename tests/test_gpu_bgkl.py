"""-m gpu parity tests of BGKLOctoMap::insert_pointcloud (src/bgkloctomap/bgkloctomap.cpp:83-268: ray segments as
training data, point-to-segment distance) through the C ABI against the CPU oracle and the reference's golden vectors."""
import numpy as np
import pytest

from conftest import golden
from util import FREE_RES, MAX_RANGE, RES, compare_leaves, key_hash, oracle_leaves_as_struct, summary

pytestmark = pytest.mark.gpu

BGKL = dict(resolution=0.1, block_depth=3, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=0.15,
            prior_A=0.001, prior_B=0.001)   # config/methods/bgkloctomap.yaml


def new_map(**kw):
    import la3dm_b200
    p = dict(BGKL)
    p.update(kw)
    return la3dm_b200.BGKLOctoMap(**p)


def test_bgkl_frontend_markers_bit_exact(scans):
    """get_training_data (bgkloctomap.cpp:285-344): re-projected hit, origin, samples walking down from l - fr."""
    from oracle.port import PortMap
    pts, org = scans["sim_structured"]
    m, o = new_map(), PortMap("bgkl")
    xy = m.training_data(pts[0], org[0], RES, FREE_RES["bgkl"], MAX_RANGE)
    want, _, _ = o.training_data(pts[0], org[0], RES, FREE_RES["bgkl"], MAX_RANGE)
    assert xy.shape == want.shape and np.array_equal(xy, want)


def test_bgkl_sequence_matches_oracle_and_golden(scans):
    from oracle.port import PortMap
    g = golden("golden_bgkl_sim_structured_seq.npz")
    pts, org = scans["sim_structured"]
    m, o = new_map(), PortMap("bgkl")
    for s in range(len(g["key_hashes"])):
        m.insert_pointcloud(pts[s], org[s], RES, FREE_RES["bgkl"], MAX_RANGE)
        o.insert_pointcloud(pts[s], org[s], RES, FREE_RES["bgkl"], MAX_RANGE)
        lv = m.leaves()
        assert key_hash(lv["block_key"], lv["depth"], lv["index"]) == str(g["key_hashes"][s]), "scan %d" % s
        compare_leaves(lv, oracle_leaves_as_struct(o.leaves()), what="bgkl scan %d" % s)
        got, want = summary(lv), g["summaries"][s]
        assert np.array_equal(got[:5], want[:5]), (s, got, want)
        assert abs(got[6] - want[6]) <= 1e-4 * want[6]
        so, sg = o.last_stats(), m.last_stats()
        for k in ("n_train", "n_data_blocks", "n_test_blocks", "voxel_visits"):
            assert so[k] == sg[k], (s, k, so[k], sg[k])
        assert so["pairs"] == sg["kernel_pairs"]
        assert abs(so["voxel_updates"] - sg["voxel_updates"]) <= max(2, so["voxel_updates"] // 5000)


def test_bgkl_other_parameters_no_range_limit(scans):
    from oracle.port import PortMap
    kw = dict(ell=0.3, sf2=0.5, var_thresh=0.3, prior_A=0.01, prior_B=0.01)
    p = dict(BGKL)
    p.update(kw)
    pts, org = scans["sim_unstructured"]
    m, o = new_map(**kw), PortMap("bgkl", p)
    for s in (4, 5):
        for mm in (m, o):
            mm.insert_pointcloud(pts[s][:2000], org[s], RES, 0.25, -1.0)
        compare_leaves(m.leaves(), oracle_leaves_as_struct(o.leaves()), what="bgkl params scan %d" % s)


@pytest.mark.parametrize("depth", [4, 5])
def test_bgkl_deeper_blocks(scans, depth):
    """config/methods/bgkloctomap_large_map.yaml uses block_depth 5 (4096 finest voxels per block): 128 leaf groups per
    (test block, neighbour) in k_bgkl_yk; k_bgkl_apply works on the 42 KB record in place (depth 4: staged)."""
    from oracle.port import PortMap
    p = dict(BGKL)
    p.update(block_depth=depth)
    pts, org = scans["sim_structured"]
    m, o = new_map(block_depth=depth), PortMap("bgkl", p)
    for s in range(2):
        m.insert_pointcloud(pts[s], org[s], RES, FREE_RES["bgkl"], MAX_RANGE)
        o.insert_pointcloud(pts[s], org[s], RES, FREE_RES["bgkl"], MAX_RANGE)
        compare_leaves(m.leaves(), oracle_leaves_as_struct(o.leaves()), what="bgkl depth %d scan %d" % (depth, s))
        so, sg = o.last_stats(), m.last_stats()
        for k in ("n_train", "n_test_blocks", "voxel_visits"):
            assert so[k] == sg[k], (s, k, so[k], sg[k])
