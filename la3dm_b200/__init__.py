"""la3dm_b200 -- B200-native per-scan Bayesian kernel occupancy update (la3dm's insert_pointcloud path).

The product is the C-ABI CUDA library la3dm_b200/lib/libla3dm_b200.so (sources in la3dm_b200/csrc, header
include/la3dm_b200.h).  This package is the thin Python mirror of the reference's map classes used by tests and
bench.py.  It never falls back to a CPU implementation.
"""
from ._lib import LIB_PATH, La3dmError, load  # noqa: F401
from .maps import (BGKLOctoMap, BGKLVOctoMap, BGKOctoMap, GPOctoMap, LEAF_DTYPE, MAP_CLASSES, NODE_DTYPE,  # noqa: F401
                   make_map)
