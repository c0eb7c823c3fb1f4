"""Host-side mirror of the reference's map classes over the C ABI.

Same names, constructor arguments (in the reference's order) and method meaning as
  la3dm::BGKOctoMap   include/bgkoctomap/bgkoctomap.h:50-58, :82-84
  la3dm::BGKLOctoMap  include/bgkloctomap/bgkloctomap.h:53-61
  la3dm::BGKLVOctoMap include/bgklvoctomap/bgklvoctomap.h:52-62
  la3dm::GPOctoMap    include/gpoctomap/gpoctomap.h:50-52
so that parity tests read like a reference node: `m = BGKOctoMap(...); m.insert_pointcloud(cloud, origin, ds, free_res,
max_range); for leaf in m.leaves(): ...`.  All work happens in the CUDA library; this file only marshals buffers.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import La3dmError, Leaf, Node, Params, ScanStats

LEAF_DTYPE = np.dtype([("block_key", "<i8"), ("depth", "<i4"), ("index", "<i4"), ("x", "<f4"), ("y", "<f4"),
                       ("z", "<f4"), ("size", "<f4"), ("a", "<f4"), ("b", "<f4"), ("prob", "<f4"), ("var", "<f4"),
                       ("state", "u1"), ("classified", "u1"), ("_pad", "u1", (6,))])
NODE_DTYPE = np.dtype([("classified", "u1"), ("_pad0", "u1", (3,)), ("a", "<f4"), ("b", "<f4"), ("state", "u1"),
                       ("_pad1", "u1", (3,))])
assert LEAF_DTYPE.itemsize == C.sizeof(Leaf) == 56 and NODE_DTYPE.itemsize == C.sizeof(Node) == 16


class _OctoMapBase:
    METHOD = None

    def __init__(self, params, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._params = params
        rc = self._lib.la3dm_create(_lib.METHODS[self.METHOD], C.byref(params), int(device), C.byref(self._h))
        if rc != 0:
            msg = self._lib.la3dm_last_error(None)
            raise La3dmError(rc, (msg or b"").decode() or self._lib.la3dm_status_string(rc).decode())

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.la3dm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise La3dmError(rc, self._lib.la3dm_last_error(self._h).decode())

    # ---- the hot path
    def insert_pointcloud(self, cloud, origin, ds_resolution, free_res=2.0, max_range=-1.0):
        """cloud: float32 [n, >=3] host array (x y z first in each row), or a CUDA torch tensor of the same shape
        (then the scan is consumed in place from device memory)."""
        o = np.ascontiguousarray(origin, dtype=np.float32)
        if hasattr(cloud, "is_cuda") and cloud.is_cuda:
            assert cloud.dtype.itemsize == 4 and cloud.dim() == 2 and cloud.stride(1) == 1 and cloud.shape[1] >= 3
            import torch
            # the map enqueues on its own non-blocking stream: order it after whatever produced the tensor
            self._check(self._lib.la3dm_stream_wait(self._h, torch.cuda.current_stream(cloud.device).cuda_stream))
            self._check(self._lib.la3dm_insert_pointcloud_device(
                self._h, cloud.data_ptr(), cloud.shape[0], cloud.stride(0) * 4, o.ctypes.data, float(ds_resolution),
                float(free_res), float(max_range)))
            return
        a = np.asarray(cloud)
        if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3 or a.strides[1] != 4 or a.strides[0] % 4:
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
        self._check(self._lib.la3dm_insert_pointcloud(self._h, a.ctypes.data, a.shape[0], a.strides[0], o.ctypes.data,
                                                      float(ds_resolution), float(free_res), float(max_range)))

    def insert_pointcloud_ingest(self, cloud, tf, prefilter_ds, origin, ds_resolution, free_res=2.0, max_range=-1.0,
                                 min_points=5):
        """cloudHandler's cloud path (bgkoctomap_server.cpp:70-86): sensor-frame cloud [n, >=3] float32, tf = 3x4 (or 4x4)
        row-major map <- sensor transform, VoxelGrid prefilter at prefilter_ds (<= 0: none), insert if > min_points."""
        a = np.asarray(cloud)
        if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3 or a.strides[1] != 4 or a.strides[0] % 4:
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(np.asarray(tf, np.float32).reshape(-1)[:12])
        o = np.ascontiguousarray(origin, dtype=np.float32)
        self._check(self._lib.la3dm_insert_pointcloud_ingest(
            self._h, a.ctypes.data, a.shape[0], a.strides[0], t.ctypes.data, float(prefilter_ds), int(min_points),
            o.ctypes.data, float(ds_resolution), float(free_res), float(max_range)))

    def insert_training_data(self, xyzy):
        """insert_training_data(xy): [n, 4] float32 host array of pre-labelled points (x y z label); BGK / GP only."""
        a = np.ascontiguousarray(xyzy, dtype=np.float32).reshape(-1, 4)
        self._check(self._lib.la3dm_insert_training_data(self._h, a.ctypes.data, a.shape[0], 16))

    def training_data(self, cloud, origin, ds_resolution, free_res=2.0, max_range=-1.0):
        """get_training_data() only: [N,7] = x0 y0 z0 x1 y1 z1 label."""
        a = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        n = C.c_size_t(0)
        args = (self._h, a.ctypes.data, a.shape[0], 12, o.ctypes.data, float(ds_resolution), float(free_res),
                float(max_range))
        self._check(self._lib.la3dm_training_data(*args, None, 0, C.byref(n)))
        out = np.zeros((n.value, 7), np.float32)
        if n.value:
            self._check(self._lib.la3dm_training_data(*args, out.ctypes.data, n.value, C.byref(n)))
        return out

    def training_rays(self):
        """BGKL / BGKLV: (rays [R,6], ray_idx [N]) of the last training_data() / insert_pointcloud() call."""
        n = C.c_size_t(0)
        self._check(self._lib.la3dm_training_rays(self._h, None, 0, C.byref(n), None, 0))
        rays = np.zeros((n.value, 6), np.float32)
        idx = np.zeros(int(self.last_stats()["n_train"]), np.int32)
        self._check(self._lib.la3dm_training_rays(self._h, rays.ctypes.data, n.value, C.byref(n), idx.ctypes.data,
                                                  idx.shape[0]))
        return rays, idx

    def last_stats(self):
        s = ScanStats()
        self._check(self._lib.la3dm_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in ScanStats._fields_}

    # ---- read side
    def get_resolution(self):
        return self._params.resolution

    def get_block_depth(self):
        return self._params.block_depth

    def get_block_size(self):
        return float(np.float32(2 ** (self._params.block_depth - 1)) * np.float32(self._params.resolution))

    def num_blocks(self):
        return int(self._lib.la3dm_num_blocks(self._h))

    def num_leaves(self):
        n = int(self._lib.la3dm_num_leaves(self._h))
        if n < 0:
            self._check(n)
        return n

    def leaves(self):
        """All leaves (begin_leaf()..end_leaf()) as a structured array sorted by (block_key, depth, index)."""
        n = C.c_size_t(0)
        self._check(self._lib.la3dm_export_leaves(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, LEAF_DTYPE)
        if n.value:
            self._check(self._lib.la3dm_export_leaves(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def touched_leaves(self, state_mask=0xFF, clear=True):
        """(block_keys, leaves) of the blocks touched since the last clearing call: the server loop's incremental mirror
        (la3dm_export_touched).  Replace everything held for the listed blocks by the returned leaves."""
        nl, nb = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.la3dm_export_touched(self._h, state_mask, None, 0, C.byref(nl), None, 0, C.byref(nb), 0))
        keys = np.zeros(nb.value, np.int64)
        out = np.zeros(nl.value, LEAF_DTYPE)
        if nb.value:
            # (with no wanted leaves at all the clearing still has to happen: a one-record dummy buffer keeps `leaves` non-NULL)
            buf = out if nl.value else np.zeros(1, LEAF_DTYPE)
            self._check(self._lib.la3dm_export_touched(self._h, state_mask, buf.ctypes.data, max(nl.value, 1), C.byref(nl),
                                                       keys.ctypes.data, nb.value, C.byref(nb), 1 if clear else 0))
        return keys, out

    def blocks(self):
        """(keys [B], nodes [B, nodes_per_block]) in the reference's Block/OcTree layout, sorted by key."""
        n = C.c_size_t(0)
        self._check(self._lib.la3dm_export_blocks(self._h, None, None, 0, C.byref(n)))
        npb = int(self._lib.la3dm_nodes_per_block(self._h))
        keys = np.zeros(n.value, np.int64)
        nodes = np.zeros((n.value, npb), NODE_DTYPE)
        if n.value:
            self._check(self._lib.la3dm_export_blocks(self._h, keys.ctypes.data, nodes.ctypes.data, n.value,
                                                      C.byref(n)))
        return keys, nodes

    def blocks_keys_only(self):
        """(keys [B] sorted, None): block keys without the node arrays (for maps too large to mirror on the host)."""
        n = C.c_size_t(0)
        self._check(self._lib.la3dm_export_blocks(self._h, None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, np.int64)
        if n.value:
            self._check(self._lib.la3dm_export_blocks(self._h, keys.ctypes.data, None, n.value, C.byref(n)))
        return keys, None

    def search(self, xyz, finest_only=False):
        """search(point3f) for a batch of points [n, 3] -> structured array of n records (LEAF_DTYPE); depth == -1 where
        the block does not exist (upstream returns a default OcTreeNode there)."""
        q = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(q), LEAF_DTYPE)
        if len(q):
            self._check(self._lib.la3dm_search(self._h, q.ctypes.data, len(q), 12, 1 if finest_only else 0,
                                               out.ctypes.data))
        return out

    def raycast(self, starts, ends, max_steps=512):
        """RayCaster(map, start, end) for a batch of rays -> (steps [n_rays, max_steps] LEAF_DTYPE, n_steps [n_rays]);
        a step outside any block has depth == -1."""
        se = np.ascontiguousarray(np.concatenate([np.asarray(starts, np.float32).reshape(-1, 3),
                                                  np.asarray(ends, np.float32).reshape(-1, 3)], 1))
        out = np.zeros((len(se), max_steps), LEAF_DTYPE)
        n = np.zeros(len(se), np.int32)
        if len(se):
            self._check(self._lib.la3dm_raycast(self._h, se.ctypes.data, len(se), int(max_steps), out.ctypes.data,
                                                n.ctypes.data))
        return out, n

    def import_blocks(self, keys, nodes):
        """Inverse of blocks(): fills an empty map from (keys [B], nodes [B, nodes_per_block])."""
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self._check(self._lib.la3dm_import_blocks(self._h, keys.ctypes.data, nodes.ctypes.data, len(keys)))

    def save(self, path):
        self._check(self._lib.la3dm_save(self._h, os.fsencode(path)))

    def load(self, path):
        self._check(self._lib.la3dm_load(self._h, os.fsencode(path)))

    def get_bbox(self):
        mn = np.zeros(3, np.float32)
        mx = np.zeros(3, np.float32)
        self._check(self._lib.la3dm_get_bbox(self._h, mn.ctypes.data, mx.ctypes.data))
        return mn, mx

    def block_to_hash_key(self, x, y, z):
        return int(self._lib.la3dm_block_to_hash_key(self._h, float(x), float(y), float(z)))

    def hash_key_to_block(self, key):
        c = np.zeros(3, np.float32)
        self._lib.la3dm_hash_key_to_block(self._h, int(key), c.ctypes.data)
        return c

    def get_extended_block(self, key):
        e = np.zeros(7, np.int64)
        self._lib.la3dm_get_extended_block(self._h, int(key), e.ctypes.data)
        return e

    # ---- multi-GPU plumbing (see include/la3dm_b200.h)
    def set_shard(self, rank, world):
        self._check(self._lib.la3dm_set_shard(self._h, int(rank), int(world)))

    def shard_rows(self):
        return int(self._lib.la3dm_shard_rows(self._h)), int(self._lib.la3dm_shard_row_bytes(self._h))

    def shard_pack(self, dev_ptr):
        self._check(self._lib.la3dm_shard_pack(self._h, int(dev_ptr)))

    def shard_unpack(self, dev_ptr):
        self._check(self._lib.la3dm_shard_unpack(self._h, int(dev_ptr)))

    def stream(self):
        return int(self._lib.la3dm_stream(self._h) or 0)

    # ---- peer replicas (BGKOctoMap): results stored straight into the other replicas' pools by the predict kernel
    def reserve_blocks(self, n):
        self._check(self._lib.la3dm_reserve_blocks(self._h, int(n)))

    def peer_local(self):
        """(pool_base, flags) device pointers of this replica, for peers living in the same process."""
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._lib.la3dm_peer_local(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def peer_ipc_export(self):
        """Two 64-byte CUDA IPC handles (pool, flags) as bytes, for peers in other processes."""
        hp, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        self._check(self._lib.la3dm_peer_ipc_export(self._h, hp, hf))
        return hp.raw, hf.raw

    def peer_ipc_open(self, handle_pool, handle_flags):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._lib.la3dm_peer_ipc_open(self._h, C.create_string_buffer(handle_pool, 64),
                                                  C.create_string_buffer(handle_flags, 64), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def peer_attach(self, world, rank, pool_bases, flags):
        P = C.c_void_p * world
        self._check(self._lib.la3dm_peer_attach(self._h, int(world), int(rank),
                                                P(*[C.c_void_p(int(v)) for v in pool_bases]),
                                                P(*[C.c_void_p(int(v)) for v in flags])))

    def peer_detach(self):
        self._check(self._lib.la3dm_peer_detach(self._h))

    def peer_set_deferred(self, on=True):
        self._check(self._lib.la3dm_peer_set_deferred(self._h, 1 if on else 0))

    def peer_sync(self):
        """Deferred mode, collective: push this rank's dirty blocks to all peers and wait for theirs."""
        self._check(self._lib.la3dm_peer_sync(self._h))


class BGKOctoMap(_OctoMapBase):
    METHOD = "bgk"

    def __init__(self, resolution=0.1, block_depth=4, sf2=1.0, ell=1.0, free_thresh=0.3, occupied_thresh=0.7,
                 var_thresh=1.0, prior_A=1.0, prior_B=1.0, device=0):
        super().__init__(Params(resolution=resolution, block_depth=block_depth, sf2=sf2, ell=ell,
                                free_thresh=free_thresh, occupied_thresh=occupied_thresh, var_thresh=var_thresh,
                                prior_A=prior_A, prior_B=prior_B), device)


class BGKLOctoMap(BGKOctoMap):
    METHOD = "bgkl"


class BGKLVOctoMap(_OctoMapBase):
    METHOD = "bgklv"

    def __init__(self, resolution=0.1, block_depth=4, sf2=1.0, ell=1.0, free_thresh=0.3, occupied_thresh=0.7,
                 var_thresh=1.0, prior_A=1.0, prior_B=1.0, original_size=True, min_W=0.1, device=0):
        super().__init__(Params(resolution=resolution, block_depth=block_depth, sf2=sf2, ell=ell,
                                free_thresh=free_thresh, occupied_thresh=occupied_thresh, var_thresh=var_thresh,
                                prior_A=prior_A, prior_B=prior_B, original_size=int(bool(original_size)),
                                min_W=min_W), device)


class GPOctoMap(_OctoMapBase):
    METHOD = "gp"

    def __init__(self, resolution=0.1, block_depth=4, sf2=1.0, ell=1.0, noise=0.01, l=100.0, min_var=0.001,
                 max_var=1000.0, max_known_var=0.02, free_thresh=0.3, occupied_thresh=0.7, device=0):
        super().__init__(Params(resolution=resolution, block_depth=block_depth, sf2=sf2, ell=ell, noise=noise, l=l,
                                min_var=min_var, max_var=max_var, max_known_var=max_known_var,
                                free_thresh=free_thresh, occupied_thresh=occupied_thresh), device)


MAP_CLASSES = {"bgk": BGKOctoMap, "bgkl": BGKLOctoMap, "bgklv": BGKLVOctoMap, "gp": GPOctoMap}


def make_map(method, params=None, device=0):
    """Build a map from a dict of reference constructor arguments (e.g. config/methods/*.yaml values)."""
    kw = dict(params or {})
    if "original_size" in kw:
        kw["original_size"] = bool(kw["original_size"])
    if "block_depth" in kw:
        kw["block_depth"] = int(kw["block_depth"])
    return MAP_CLASSES[method](device=device, **kw)
