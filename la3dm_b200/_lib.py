"""ctypes binding of la3dm_b200/lib/libla3dm_b200.so -- the C ABI declared in include/la3dm_b200.h.

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libla3dm_b200.so")

METHODS = {"bgk": 0, "bgkl": 1, "bgklv": 2, "gp": 3}

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_EXTENT, ERR_NOMEM, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6


class Params(C.Structure):
    _fields_ = [("resolution", C.c_float), ("block_depth", C.c_int32), ("sf2", C.c_float), ("ell", C.c_float),
                ("free_thresh", C.c_float), ("occupied_thresh", C.c_float), ("var_thresh", C.c_float),
                ("prior_A", C.c_float), ("prior_B", C.c_float), ("original_size", C.c_int32), ("min_W", C.c_float),
                ("noise", C.c_float), ("l", C.c_float), ("min_var", C.c_float), ("max_var", C.c_float),
                ("max_known_var", C.c_float)]


class Node(C.Structure):
    _fields_ = [("classified", C.c_uint8), ("_pad0", C.c_uint8 * 3), ("a", C.c_float), ("b", C.c_float),
                ("state", C.c_uint8), ("_pad1", C.c_uint8 * 3)]


class Leaf(C.Structure):
    _fields_ = [("block_key", C.c_int64), ("depth", C.c_int32), ("index", C.c_int32), ("x", C.c_float),
                ("y", C.c_float), ("z", C.c_float), ("size", C.c_float), ("a", C.c_float), ("b", C.c_float),
                ("prob", C.c_float), ("var", C.c_float), ("state", C.c_uint8), ("classified", C.c_uint8),
                ("_pad", C.c_uint8 * 6)]


class ScanStats(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_hits", C.c_int64), ("n_train", C.c_int64), ("n_data_blocks", C.c_int64),
                ("n_test_blocks", C.c_int64), ("voxel_visits", C.c_int64), ("voxel_updates", C.c_int64),
                ("kernel_pairs", C.c_int64), ("n_blocks_total", C.c_int64), ("new_blocks", C.c_int64),
                ("kernel_launches", C.c_int32), ("grid_irregular", C.c_int32), ("device_ms", C.c_float),
                ("predict_ms", C.c_float), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("replays", C.c_int32), ("graph_captures", C.c_int32)]


# every symbol include/la3dm_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "la3dm_create": (C.c_int, [C.c_int, C.POINTER(Params), C.c_int, C.POINTER(_P)]),
    "la3dm_destroy": (C.c_int, [_P]),
    "la3dm_last_error": (C.c_char_p, [_P]),
    "la3dm_status_string": (C.c_char_p, [C.c_int]),
    "la3dm_abi_version": (C.c_int, []),
    "la3dm_insert_pointcloud": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, C.c_float, C.c_float, C.c_float]),
    "la3dm_insert_pointcloud_device": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, C.c_float, C.c_float,
                                                 C.c_float]),
    "la3dm_insert_pointcloud_ingest": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, C.c_float, C.c_int, _P, C.c_float,
                                                C.c_float, C.c_float]),
    "la3dm_insert_training_data": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "la3dm_insert_training_data_device": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "la3dm_training_data": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, C.c_float, C.c_float, C.c_float, _P,
                                      C.c_size_t, C.POINTER(C.c_size_t)]),
    "la3dm_training_rays": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t), _P, C.c_size_t]),
    "la3dm_last_stats": (C.c_int, [_P, C.POINTER(ScanStats)]),
    "la3dm_num_blocks": (C.c_int64, [_P]),
    "la3dm_nodes_per_block": (C.c_int32, [_P]),
    "la3dm_export_blocks": (C.c_int, [_P, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "la3dm_num_leaves": (C.c_int64, [_P]),
    "la3dm_export_leaves": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "la3dm_export_touched": (C.c_int, [_P, C.c_uint, _P, C.c_size_t, C.POINTER(C.c_size_t), _P, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.c_int]),
    "la3dm_search": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.c_int, _P]),
    "la3dm_raycast": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, _P]),
    "la3dm_import_blocks": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "la3dm_save": (C.c_int, [_P, C.c_char_p]),
    "la3dm_load": (C.c_int, [_P, C.c_char_p]),
    "la3dm_get_bbox": (C.c_int, [_P, _P, _P]),
    "la3dm_block_to_hash_key": (C.c_int64, [_P, C.c_float, C.c_float, C.c_float]),
    "la3dm_hash_key_to_block": (None, [_P, C.c_int64, _P]),
    "la3dm_get_extended_block": (None, [_P, C.c_int64, _P]),
    "la3dm_set_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "la3dm_shard_row_bytes": (C.c_int64, [_P]),
    "la3dm_shard_rows": (C.c_int64, [_P]),
    "la3dm_shard_pack": (C.c_int, [_P, _P]),
    "la3dm_shard_unpack": (C.c_int, [_P, _P]),
    "la3dm_reserve_blocks": (C.c_int, [_P, C.c_size_t]),
    "la3dm_peer_local": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "la3dm_peer_ipc_export": (C.c_int, [_P, _P, _P]),
    "la3dm_peer_ipc_open": (C.c_int, [_P, _P, _P, C.POINTER(_P), C.POINTER(_P)]),
    "la3dm_peer_attach": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    "la3dm_peer_detach": (C.c_int, [_P]),
    "la3dm_peer_set_deferred": (C.c_int, [_P, C.c_int]),
    "la3dm_peer_sync": (C.c_int, [_P]),
    "la3dm_stream": (_P, [_P]),
    "la3dm_stream_wait": (C.c_int, [_P, _P]),
    "la3dm_bench_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """dlopen the product library (raises if it has not been built: run `python -c 'import __graft_entry__ as g;
    g.build()'` or `make -C la3dm_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built (make -C la3dm_b200/csrc); la3dm_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class La3dmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("la3dm_b200: %s (status %d)" % (msg, status))
        self.status = status
