"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libla3dm_ref_<method>[_fast].so.

Those libraries are the reference's own map sources (compiled in place from /root/reference by oracle/Makefile against
the stand-in PCL / inference headers in oracle/standins) behind the small C driver oracle/ref_driver.cpp.  Only tests/,
tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's reference / cpu_baseline arm may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = ("bgk", "bgkl", "bgklv", "gp")

# parameter vectors in the order of the reference constructors (see ref_driver.cpp:ref_create)
DEFAULT_PARAMS = {
    # config/methods/bgkoctomap.yaml
    "bgk": dict(resolution=0.1, block_depth=3, sf2=1.0, ell=0.2, free_thresh=0.3, occupied_thresh=0.7,
                var_thresh=100.0, prior_A=0.001, prior_B=0.001),
    # config/methods/bgkloctomap.yaml
    "bgkl": dict(resolution=0.1, block_depth=3, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7,
                 var_thresh=0.15, prior_A=0.001, prior_B=0.001),
    # config/methods/bgklvoctomap.yaml
    "bgklv": dict(resolution=0.1, block_depth=5, sf2=0.1, ell=0.2, free_thresh=0.3, occupied_thresh=0.7,
                  var_thresh=0.2, prior_A=0.001, prior_B=0.001, original_size=0.0, min_W=0.001),
    # config/methods/gpoctomap.yaml
    "gp": dict(resolution=0.1, block_depth=3, sf2=1.0, ell=1.0, noise=0.01, l=100.0, min_var=0.001,
               max_var=1000.0, max_known_var=0.02, free_thresh=0.3, occupied_thresh=0.7),
}


def lib_path(method, fast=False):
    return os.path.join(_HERE, "_ref", "libla3dm_ref_%s%s.so" % (method, "_fast" if fast else ""))


def available(method="bgk", fast=False):
    return os.path.exists(lib_path(method, fast))


class RefMap:
    """One reference map object (la3dm::BGKOctoMap | BGKLOctoMap | BGKLVOctoMap | GPOctoMap)."""

    def __init__(self, method="bgk", params=None, fast=False, threads=None):
        assert method in METHODS
        self.method = method
        p = dict(DEFAULT_PARAMS[method])
        if params:
            p.update(params)
        self.params = p
        path = lib_path(method, fast)
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `make -C oracle ref` where /root/reference exists" % path)
        # RTLD_LOCAL: the four libraries define the same la3dm:: symbols with different layouts
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_void_p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_insert_pointcloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                            C.c_float]
        L.ref_insert_training_data.restype = C.c_int
        L.ref_insert_training_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ref_num_blocks.restype = C.c_int64
        L.ref_num_blocks.argtypes = [C.c_void_p]
        L.ref_num_leaves.restype = C.c_int64
        L.ref_num_leaves.argtypes = [C.c_void_p]
        L.ref_dump_leaves.argtypes = [C.c_void_p] * 9
        L.ref_get_bbox.argtypes = [C.c_void_p] * 3
        L.ref_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_raycast.restype = C.c_int64
        L.ref_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.ref_training_data.restype = C.c_int64
        L.ref_training_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p]
        L.ref_training_data_copy.argtypes = [C.c_void_p] * 3
        L.ref_block_to_hash_key.restype = C.c_int64
        L.ref_block_to_hash_key.argtypes = [C.c_float] * 3
        L.ref_hash_key_to_block.argtypes = [C.c_int64, C.c_void_p]
        L.ref_extended_block.argtypes = [C.c_int64, C.c_void_p]
        L.ref_key_loc.restype = C.c_int
        L.ref_key_loc.argtypes = [C.c_int, C.c_int, C.c_void_p]
        if threads is not None:
            L.ref_set_threads(int(threads))
        vec = np.asarray(list(p.values()), dtype=np.float32)
        self.h = L.ref_create(vec.ctypes.data, len(vec))
        assert self.h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, n):
        self.lib.ref_set_threads(int(n))

    def max_threads(self):
        return int(self.lib.ref_max_threads())

    def insert_pointcloud(self, xyz, origin, ds_resolution, free_res=2.0, max_range=-1.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        self.lib.ref_insert_pointcloud(self.h, xyz.ctypes.data, xyz.shape[0], o.ctypes.data,
                                       float(ds_resolution), float(free_res), float(max_range))

    def insert_training_data(self, xyzy):
        """insert_training_data(xy), BGK / GP; every test block must already exist (upstream null dereference)."""
        a = np.ascontiguousarray(xyzy, dtype=np.float32).reshape(-1, 4)
        rc = self.lib.ref_insert_training_data(self.h, a.ctypes.data, a.shape[0])
        if rc != 0:
            raise NotImplementedError("insert_training_data: BGK / GP only")

    def num_blocks(self):
        return int(self.lib.ref_num_blocks(self.h))

    def leaves(self):
        """dict of arrays sorted by (block_key, depth, index)."""
        n = int(self.lib.ref_num_leaves(self.h))
        out = dict(block_key=np.zeros(n, np.int64), depth=np.zeros(n, np.int32), index=np.zeros(n, np.int32),
                   loc_size=np.zeros((n, 4), np.float32), ab=np.zeros((n, 2), np.float32),
                   state=np.zeros(n, np.uint8), classified=np.zeros(n, np.uint8),
                   prob_var=np.zeros((n, 2), np.float32))
        self.lib.ref_dump_leaves(self.h, *[out[k].ctypes.data for k in
                                           ("block_key", "depth", "index", "loc_size", "ab", "state", "classified",
                                            "prob_var")])
        order = np.lexsort((out["index"], out["depth"], out["block_key"]))
        return {k: v[order] for k, v in out.items()}

    def search(self, xyz):
        """search(x, y, z) for n points -> (ab [n, 2], state [n], classified [n]) of the node upstream returns."""
        q = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        ab = np.zeros((len(q), 2), np.float32)
        st = np.zeros(len(q), np.uint8)
        cl = np.zeros(len(q), np.uint8)
        self.lib.ref_search(self.h, q.ctypes.data, len(q), ab.ctypes.data, st.ctypes.data, cl.ctypes.data)
        return ab, st, cl

    def raycast(self, start, end, max_steps=512):
        """RayCaster(map, start, end): dict of per-step arrays (p [n,3], block_key, node_key, valid, ab [n,2], state)."""
        a, b = np.ascontiguousarray(start, np.float32), np.ascontiguousarray(end, np.float32)
        p = np.zeros((max_steps, 3), np.float32)
        bk, nk = np.zeros(max_steps, np.int64), np.zeros(max_steps, np.int64)
        ok, st = np.zeros(max_steps, np.uint8), np.zeros(max_steps, np.uint8)
        ab = np.zeros((max_steps, 2), np.float32)
        n = int(self.lib.ref_raycast(self.h, a.ctypes.data, b.ctypes.data, max_steps, p.ctypes.data, bk.ctypes.data,
                                     nk.ctypes.data, ok.ctypes.data, ab.ctypes.data, st.ctypes.data))
        return dict(p=p[:n], block_key=bk[:n], node_key=nk[:n], valid=ok[:n], ab=ab[:n], state=st[:n])

    def get_bbox(self):
        mn = np.zeros(3, np.float32)
        mx = np.zeros(3, np.float32)
        self.lib.ref_get_bbox(self.h, mn.ctypes.data, mx.ctypes.data)
        return mn, mx

    def training_data(self, xyz, origin, ds_resolution, free_res, max_range):
        """(xy7 [N,7] = x0 y0 z0 x1 y1 z1 label, ray_idx [N], rays [R,6])"""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        nr = C.c_int64(0)
        n = int(self.lib.ref_training_data(self.h, xyz.ctypes.data, xyz.shape[0], o.ctypes.data,
                                           float(ds_resolution), float(free_res), float(max_range), C.byref(nr)))
        xy = np.zeros((n, 7), np.float32)
        ri = np.zeros(n, np.int32)
        rays = np.zeros((nr.value, 6), np.float32)
        self.lib.ref_training_data_copy(xy.ctypes.data, ri.ctypes.data, rays.ctypes.data)
        return xy, ri, rays

    def block_to_hash_key(self, x, y, z):
        return int(self.lib.ref_block_to_hash_key(float(x), float(y), float(z)))

    def hash_key_to_block(self, key):
        c = np.zeros(3, np.float32)
        self.lib.ref_hash_key_to_block(int(key), c.ctypes.data)
        return c

    def extended_block(self, key):
        e = np.zeros(7, np.int64)
        self.lib.ref_extended_block(int(key), e.ctypes.data)
        return e

    def key_loc(self, depth, index):
        c = np.zeros(3, np.float32)
        ok = self.lib.ref_key_loc(int(depth), int(index), c.ctypes.data)
        return c if ok else None


def read_pcd(path):
    """Binary PCD (FIELDS x y z intensity, float32) + sensor origin from VIEWPOINT, as the static nodes load it
    (src/bgkoctomap/bgkoctomap_static_node.cpp:7-16)."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr_end = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    hdr = raw[:hdr_end].decode("ascii", "replace").splitlines()
    meta = {l.split()[0]: l.split()[1:] for l in hdr if l and not l.startswith("#")}
    n = int(meta["POINTS"][0])
    nf = len(meta["FIELDS"])
    pts = np.frombuffer(raw, dtype=np.float32, count=n * nf, offset=hdr_end).reshape(n, nf)[:, :3].copy()
    origin = np.array([float(v) for v in meta["VIEWPOINT"][:3]], dtype=np.float32)
    return pts, origin
