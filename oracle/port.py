"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libla3dm_oracle.so (our CPU restatement, la3dm_oracle.cpp).

Same Python surface as oracle/ref.py:RefMap so that tests can swap the two.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this; the product package la3dm_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .ref import DEFAULT_PARAMS, METHODS

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libla3dm_oracle.so")
METHOD_ID = {"bgk": 0, "bgkl": 1, "bgklv": 2, "gp": 3}


def build(force=False):
    src = os.path.join(_HERE, "la3dm_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_insert_pointcloud.restype = C.c_int
        L.orc_insert_pointcloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                            C.c_float]
        L.orc_last_stats.argtypes = [C.c_void_p, C.c_void_p]
        for f in (L.orc_num_blocks, L.orc_num_leaves):
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        L.orc_dump_leaves.argtypes = [C.c_void_p] * 9
        L.orc_training_data.restype = C.c_int64
        L.orc_training_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p]
        L.orc_training_data_copy.argtypes = [C.c_void_p] * 4
        L.orc_voxel_grid.restype = C.c_int64
        L.orc_voxel_grid.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        L.orc_block_to_hash_key.restype = C.c_int64
        L.orc_block_to_hash_key.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.orc_extended_block.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_key_loc.restype = C.c_int
        L.orc_key_loc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def voxel_grid(xyz, leaf):
    L = _load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(xyz)
    n = int(L.orc_voxel_grid(xyz.ctypes.data, xyz.shape[0], float(leaf), out.ctypes.data))
    return out[:n].copy()


class PortMap:
    STAT_NAMES = ("n_train", "n_data_blocks", "n_test_blocks", "voxel_visits", "voxel_updates", "pairs")

    def __init__(self, method="bgk", params=None):
        assert method in METHODS
        self.method = method
        p = dict(DEFAULT_PARAMS[method])
        if params:
            p.update(params)
        self.params = p
        self.lib = _load()
        vec = np.asarray(list(p.values()), dtype=np.float32)
        self.h = self.lib.orc_create(METHOD_ID[method], vec.ctypes.data, len(vec))
        if not self.h:
            raise NotImplementedError("port oracle: method %s not restated yet" % method)

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert_pointcloud(self, xyz, origin, ds_resolution, free_res=2.0, max_range=-1.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        rc = self.lib.orc_insert_pointcloud(self.h, xyz.ctypes.data, xyz.shape[0], o.ctypes.data,
                                            float(ds_resolution), float(free_res), float(max_range))
        if rc != 0:
            raise RuntimeError("orc_insert_pointcloud -> %d" % rc)

    def last_stats(self):
        s = np.zeros(6, np.int64)
        self.lib.orc_last_stats(self.h, s.ctypes.data)
        return dict(zip(self.STAT_NAMES, (int(v) for v in s)))

    def num_blocks(self):
        return int(self.lib.orc_num_blocks(self.h))

    def leaves(self):
        n = int(self.lib.orc_num_leaves(self.h))
        out = dict(block_key=np.zeros(n, np.int64), depth=np.zeros(n, np.int32), index=np.zeros(n, np.int32),
                   loc_size=np.zeros((n, 4), np.float32), ab=np.zeros((n, 2), np.float32),
                   state=np.zeros(n, np.uint8), classified=np.zeros(n, np.uint8),
                   prob_var=np.zeros((n, 2), np.float32))
        self.lib.orc_dump_leaves(self.h, *[out[k].ctypes.data for k in
                                           ("block_key", "depth", "index", "loc_size", "ab", "state", "classified",
                                            "prob_var")])
        order = np.lexsort((out["index"], out["depth"], out["block_key"]))
        return {k: v[order] for k, v in out.items()}

    def training_data(self, xyz, origin, ds_resolution, free_res, max_range):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        nr = C.c_int64(0)
        n = int(self.lib.orc_training_data(self.h, xyz.ctypes.data, xyz.shape[0], o.ctypes.data,
                                           float(ds_resolution), float(free_res), float(max_range), C.byref(nr)))
        xy = np.zeros((n, 7), np.float32)
        ri = np.zeros(n, np.int32)
        rays = np.zeros((nr.value, 6), np.float32)
        self.lib.orc_training_data_copy(self.h, xy.ctypes.data, ri.ctypes.data, rays.ctypes.data)
        return xy, ri, rays

    def block_to_hash_key(self, x, y, z):
        return int(self.lib.orc_block_to_hash_key(self.h, float(x), float(y), float(z)))

    def extended_block(self, key):
        e = np.zeros(7, np.int64)
        self.lib.orc_extended_block(self.h, int(key), e.ctypes.data)
        return e

    def key_loc(self, depth, index):
        c = np.zeros(3, np.float32)
        ok = self.lib.orc_key_loc(self.h, int(depth), int(index), c.ctypes.data)
        return c if ok else None
