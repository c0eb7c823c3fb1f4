"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/la3dm_b200.h declares, the
struct layouts agree with the header, and the library refuses to work without a GPU (there is no CPU fallback).
No compute calls here."""
import ctypes as C
import os
import re

import pytest

from la3dm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "la3dm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(la3dm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    # the ctypes table binds exactly the header's functions
    assert sorted(_lib.SYMBOLS) == names


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Node) == 16            # reference Occupancy node (bgkoctree_node.h:76-81)
    assert C.sizeof(_lib.Leaf) == 56
    assert C.sizeof(_lib.Params) == 16 * 4
    assert _lib.Leaf.block_key.offset == 0 and _lib.Leaf.x.offset == 16 and _lib.Leaf.state.offset == 48
    assert C.sizeof(_lib.ScanStats) == 10 * 8 + 4 * 4 + 2 * 8 + 2 * 4


def test_abi_version_and_status_strings():
    lib = _lib.load()
    assert lib.la3dm_abi_version() == 1
    assert lib.la3dm_status_string(0) == b"ok"
    assert lib.la3dm_status_string(_lib.ERR_NO_DEVICE) == b"no CUDA device"
    assert lib.la3dm_status_string(-99) == b"unknown status"


def test_null_arguments_are_rejected():
    lib = _lib.load()
    assert lib.la3dm_create(0, None, 0, None) == _lib.ERR_INVALID
    assert lib.la3dm_destroy(None) == _lib.ERR_INVALID
    assert lib.la3dm_num_blocks(None) == -1
    assert lib.la3dm_insert_pointcloud(None, None, 0, 12, None, 0.1, 0.5, -1.0) == _lib.ERR_INVALID


def test_no_cpu_fallback():
    """Without a CUDA device the constructor must fail loudly (LA3DM_ERR_NO_DEVICE), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import la3dm_b200
    with pytest.raises(la3dm_b200.La3dmError) as e:
        la3dm_b200.BGKOctoMap(resolution=0.1, block_depth=3)
    assert e.value.status == _lib.ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "la3dm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle[/.]|la3dm_oracle|la3dm_ref", txt), \
                    "%s reaches into oracle/" % f
