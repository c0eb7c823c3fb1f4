"""-m gpu tests of the ingest row: the live node's steps in front of insert_pointcloud --
pcl_ros::transformPointCloud into the map frame and the pcl::VoxelGrid prefilter (src/bgkoctomap/bgkoctomap_server.cpp:
70-86) -- fused into the GPU front-end behind la3dm_insert_pointcloud_ingest.  Checker: the same two steps restated in
numpy / the oracle's VoxelGrid (oracle/port.py:voxel_grid), then the oracle's insert_pointcloud."""
import numpy as np
import pytest

from util import FREE_RES, MAX_RANGE, RES, compare_leaves, gp_compare, oracle_leaves_as_struct

pytestmark = pytest.mark.gpu

BGK = dict(resolution=0.1, block_depth=3, sf2=1.0, ell=0.2, free_thresh=0.3, occupied_thresh=0.7, var_thresh=100.0,
           prior_A=0.001, prior_B=0.001)


def transform_like_pcl(tf, xyz):
    """pcl::transformPointCloud (pcl::detail::Transformer, PCL >= 1.10): (m0 x + m1 y) + (m2 z + m3) per row in fp32."""
    m = np.asarray(tf, np.float32).reshape(3, 4)
    x, y, z = (xyz[:, k].astype(np.float32) for k in range(3))
    out = [(m[r, 0] * x + m[r, 1] * y) + (m[r, 2] * z + m[r, 3]) for r in range(3)]
    return np.stack(out, 1).astype(np.float32)


def pose(yaw, pitch, t):
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    return np.concatenate([R, np.asarray(t, np.float64).reshape(3, 1)], 1).astype(np.float32)


def test_ingest_matches_transform_prefilter_insert(scans):
    import la3dm_b200
    from oracle import port
    pts, org = scans["sim_unstructured"]
    m, o = la3dm_b200.BGKOctoMap(**BGK), port.PortMap("bgk")
    for s in range(4):
        # the shipped scans are in the map frame: move them into a made-up sensor frame, hand that cloud + the pose over
        T = pose(0.3 * s - 0.4, 0.05 * s, org[s])
        R, t = T[:, :3].astype(np.float64), T[:, 3].astype(np.float64)
        sensor = ((pts[s].astype(np.float64) - t) @ R).astype(np.float32)          # R^T (p - t)
        pre_ds = 0.15 if s != 2 else -1.0                                            # scan 2: no prefilter (-LV server)
        m.insert_pointcloud_ingest(sensor, T, pre_ds, T[:, 3], RES, FREE_RES["bgk"], MAX_RANGE, min_points=5)
        world = transform_like_pcl(T, sensor)
        filt = port.voxel_grid(world, pre_ds) if pre_ds > 0 else world
        assert len(filt) > 5
        o.insert_pointcloud(filt, T[:, 3], RES, FREE_RES["bgk"], MAX_RANGE)
        assert m.last_stats()["n_train"] == o.last_stats()["n_train"], s
        compare_leaves(m.leaves(), oracle_leaves_as_struct(o.leaves()), what="ingest scan %d" % s)
    # too few points after the prefilter: the call is a no-op (bgkoctomap_server.cpp:84)
    before = m.leaves().tobytes()
    few = pts[0][:40] * np.float32(0.01)                       # 40 points inside one 0.5 m prefilter voxel or two
    m.insert_pointcloud_ingest(few, pose(0, 0, [0, 0, 0]), 0.5, np.zeros(3, np.float32), RES, 0.5, MAX_RANGE, min_points=5)
    assert m.last_stats()["n_train"] == 0 and m.leaves().tobytes() == before
    # identity transform, no prefilter == plain insert_pointcloud
    a, b = la3dm_b200.BGKOctoMap(**BGK), la3dm_b200.BGKOctoMap(**BGK)
    a.insert_pointcloud_ingest(pts[0], pose(0, 0, [0, 0, 0]), -1.0, org[0], RES, 0.5, MAX_RANGE, min_points=0)
    b.insert_pointcloud(pts[0], org[0], RES, 0.5, MAX_RANGE)
    assert a.leaves().tobytes() == b.leaves().tobytes()


def test_ingest_gp_and_bgkl(scans):
    """the other map classes take the same path (their servers share the handler)."""
    import la3dm_b200
    from oracle import port, ref
    pts, org = scans["sim_structured"]
    for method in ("gp", "bgkl"):
        kw = dict(ref.DEFAULT_PARAMS[method])
        kw["block_depth"] = int(kw["block_depth"])
        m, o = la3dm_b200.maps.make_map(method, kw), port.PortMap(method)
        T = pose(-0.7, 0.02, org[1])
        R, t = T[:, :3].astype(np.float64), T[:, 3].astype(np.float64)
        sensor = ((pts[1].astype(np.float64) - t) @ R).astype(np.float32)
        m.insert_pointcloud_ingest(sensor, T, 0.12, T[:, 3], RES, FREE_RES[method], MAX_RANGE)
        filt = port.voxel_grid(transform_like_pcl(T, sensor), 0.12)
        o.insert_pointcloud(filt, T[:, 3], RES, FREE_RES[method], MAX_RANGE)
        assert m.last_stats()["n_train"] == o.last_stats()["n_train"], method
        (gp_compare if method == "gp" else compare_leaves)(m.leaves(), oracle_leaves_as_struct(o.leaves()), what="ingest " + method)
