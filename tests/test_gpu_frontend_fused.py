"""-m gpu: the sort-free cooperative front-end (la3dm_b200/csrc/frontend_fused.cu) against the sort-based pipeline it
replaces (frontend.cu + binning.cu, LA3DM_LEGACY_FRONTEND=1), on the device, bit for bit: training sets, unit counts and
every exported leaf.  The legacy pipeline itself is pinned on the reference's golden vectors (test_gpu_bgk.py, which
runs on the fused path by default as well); this file covers what those scans do not reach: spans handed to a warp /
a CTA, points on block boundaries, samples inside the sensor's own voxel, the fall-back for spans that are too long."""
import os

import numpy as np
import pytest

from util import FREE_RES, MAX_RANGE, RES

pytestmark = pytest.mark.gpu

BGK = dict(resolution=0.1, block_depth=3, sf2=1.0, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=100.0,
           prior_A=0.001, prior_B=0.001)
GP = dict(resolution=0.1, block_depth=3, sf2=1.0, ell=1.0, noise=0.01, l=100.0, min_var=0.001, max_var=1000.0,
          max_known_var=0.02, free_thresh=0.3, occupied_thresh=0.7)


def make(method, legacy, **kw):
    """The front-end flavour is chosen when the map is created (la3dm_create reads the environment)."""
    import la3dm_b200
    old = os.environ.pop("LA3DM_LEGACY_FRONTEND", None)
    if legacy:
        os.environ["LA3DM_LEGACY_FRONTEND"] = "1"
    try:
        if method == "gp":
            p = dict(GP); p.update(kw)
            return la3dm_b200.GPOctoMap(**p)
        p = dict(BGK); p.update(kw)
        return la3dm_b200.BGKOctoMap(**p)
    finally:
        os.environ.pop("LA3DM_LEGACY_FRONTEND", None)
        if old is not None:
            os.environ["LA3DM_LEGACY_FRONTEND"] = old


UNIT_KEYS = ("n_hits", "n_train", "n_data_blocks", "n_test_blocks", "voxel_visits", "voxel_updates", "kernel_pairs",
             "new_blocks")


def same_scan(a, b, pts, org, ds, fr, mr, what):
    """insert into both maps; unit counts and all leaves must be identical (bytes)."""
    a.insert_pointcloud(pts, org, ds, fr, mr)
    b.insert_pointcloud(pts, org, ds, fr, mr)
    sa, sb = a.last_stats(), b.last_stats()
    for k in UNIT_KEYS:
        assert sa[k] == sb[k], (what, k, sa[k], sb[k])
    la, lb = a.leaves(), b.leaves()
    assert la.shape == lb.shape, what
    assert la.tobytes() == lb.tobytes(), what
    return sa, sb


def test_fused_equals_legacy_on_shipped_and_synthetic_scans(scans):
    from la3dm_b200.synthetic import make_sequence
    pts, org = scans["sim_unstructured"]
    f, l = make("bgk", False), make("bgk", True)
    for s in range(6):
        xa = f.training_data(pts[s], org[s], RES, FREE_RES["bgk"], MAX_RANGE)
        xb = l.training_data(pts[s], org[s], RES, FREE_RES["bgk"], MAX_RANGE)
        assert np.array_equal(xa, xb), s
        sa, sb = same_scan(f, l, pts[s], org[s], RES, FREE_RES["bgk"], MAX_RANGE, "sim_unstructured %d" % s)
    assert sa["kernel_launches"] <= 12 < sb["kernel_launches"], (sa["kernel_launches"], sb["kernel_launches"])
    assert sa["replays"] == 0
    # the bench workload: long spans next to the sensor (hundreds of samples per voxel at d = free_res)
    big, borg = make_sequence(3, 65536, 50.0, seed=1)
    f, l = make("bgk", False), make("bgk", True)
    for s in range(3):
        xa = f.training_data(big[s], borg[s], RES, 0.5, -1.0)
        xb = l.training_data(big[s], borg[s], RES, 0.5, -1.0)
        assert np.array_equal(xa, xb), s
        sa, sb = same_scan(f, l, big[s], borg[s], RES, 0.5, -1.0, "synthetic 64k %d" % s)
    assert sa["kernel_launches"] <= 12 and sa["replays"] == 0, sa


def test_fused_equals_legacy_gp_and_training_data(scans):
    pts, org = scans["sim_unstructured"]
    f, l = make("gp", False), make("gp", True)
    for s in range(4):
        same_scan(f, l, pts[s], org[s], RES, FREE_RES["gp"], MAX_RANGE, "gp %d" % s)
    # insert_training_data: the binning / plan stages alone
    f, l = make("bgk", False), make("bgk", True)
    xy = l.training_data(pts[0], org[0], RES, FREE_RES["bgk"], MAX_RANGE)[:, [0, 1, 2, 6]].copy()
    for m in (f, l):
        m.insert_training_data(xy)
    assert f.leaves().tobytes() == l.leaves().tobytes()
    assert f.last_stats()["n_test_blocks"] == l.last_stats()["n_test_blocks"]


def test_points_on_block_boundaries_and_identity_downsampling():
    """ds_resolution < 0: downsample() is the identity, so the training points are the cloud itself -- put them ON the
    block faces (closed boxes: a point belongs to every block that touches it, rtree.h:1519-1532) and edges.  (1 300
    points: the sensor's block holds one origin copy per hit, and more than 2 048 entries in a block would send the scan
    to the legacy pipeline -- see the last test.)"""
    rng = np.random.default_rng(5)
    g = (np.arange(-10, 11) * 0.4 + 0.2).astype(np.float32)        # block faces of a 0.4 m block grid
    faces = np.stack([rng.choice(g, 300), rng.uniform(-4, 4, 300), rng.uniform(0, 2, 300)], 1)
    edges = np.stack([rng.choice(g, 150), rng.choice(g, 150), rng.uniform(0, 2, 150)], 1)
    corners = np.stack([rng.choice(g, 50), rng.choice(g, 50), rng.choice(g[8:14], 50)], 1)
    pts = np.concatenate([faces, edges, corners, rng.uniform(-4, 4, (800, 3))]).astype(np.float32)
    org = np.array([0.2, 0.2, 1.0], np.float32)                    # the origin on an edge as well
    f, l = make("bgk", False), make("bgk", True)
    sa, sb = same_scan(f, l, pts, org, -1.0, 0.4, -1.0, "boundaries")
    assert sa["kernel_launches"] <= 12          # (capacity replays are fine; still on the fused path)
    assert sa["n_train"] > len(pts)


def test_samples_inside_the_sensor_voxel():
    """A small free_resolution puts beam samples into the origin's own voxel, between the origin copies: the centroid
    of that voxel is a sequential sum in push order (bgkoctomap.cpp:404, :451-457; pcl CentroidPoint)."""
    rng = np.random.default_rng(11)
    d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (np.array([1.03, -0.52, 0.77]) + d * rng.uniform(0.02, 3.0, (4000, 1))).astype(np.float32)
    org = np.array([1.03, -0.52, 0.77], np.float32)
    f, l = make("bgk", False), make("bgk", True)
    xa = f.training_data(pts, org, RES, 0.03, -1.0)
    xb = l.training_data(pts, org, RES, 0.03, -1.0)
    assert np.array_equal(xa, xb)
    same_scan(f, l, pts, org, RES, 0.03, -1.0, "sensor voxel")


def test_span_too_long_replays_on_the_legacy_pipeline(scans):
    """More than 2 048 points in one voxel: the fused front-end raises its overflow bit before the map is touched, the
    host replays the scan on the sort-based pipeline and stays there; results equal the legacy map's."""
    rng = np.random.default_rng(2)
    pts, org = scans["sim_structured"]
    blob = (np.array([2.03, 1.02, 0.51]) + rng.uniform(0, 0.04, (5000, 3))).astype(np.float32)    # one 0.1 m voxel
    cloud = np.concatenate([pts[0], blob]).astype(np.float32)
    f, l = make("bgk", False), make("bgk", True)
    same_scan(f, l, pts[1], org[1], RES, 0.5, MAX_RANGE, "before")
    sa, sb = same_scan(f, l, cloud, org[0], RES, 0.5, MAX_RANGE, "blob")
    assert sa["replays"] >= 1 and sa["kernel_launches"] > 12, sa
    sa, sb = same_scan(f, l, pts[2], org[2], RES, 0.5, MAX_RANGE, "after")
    assert sa["kernel_launches"] > 12
