"""Multi-GPU plumbing of one scan: the test blocks of a scan are dealt round-robin over ranks
(test block t belongs to rank t % world; the CUDA predict kernel skips the others), every rank packs the node arrays of
ITS blocks into fixed-size rows, ONE all-gather moves the rows (NCCL over NVLink on the GPU box, gloo in the CPU
tests), and every rank scatters the peers' rows into its replica.

Only index arithmetic and the collective live here; pack / unpack are the CUDA kernels behind la3dm_shard_pack /
la3dm_shard_unpack (la3dm_b200/csrc/shard.cu).
"""


def rows_per_rank(n_test_blocks, world):
    """la3dm_shard_rows(): every rank contributes the same number of rows (the tail is padding)."""
    return (int(n_test_blocks) + world - 1) // world


def owner_of(t, world):
    return t % world


def row_of(t, world):
    """Row of test block t inside its owner's packed buffer."""
    return t // world


def test_block_of(rank, row, world):
    """Inverse of (owner_of, row_of); may be >= n_test_blocks for padding rows."""
    return row * world + rank


def exchange(map_, world, all_gather, alloc):
    """Runs the per-scan exchange for `map_` (an object with shard_rows / shard_pack / shard_unpack).

    alloc(nbytes) -> (buffer object, address);  all_gather(out_buffer, in_buffer) performs the collective.
    Returns the number of collectives issued (0 when the scan had no test blocks)."""
    rows, row_bytes = map_.shard_rows()
    if rows == 0 or world == 1:
        return 0
    mine, mine_ptr = alloc(rows * row_bytes)
    allr, all_ptr = alloc(world * rows * row_bytes)
    map_.shard_pack(mine_ptr)
    all_gather(allr, mine)
    map_.shard_unpack(all_ptr)
    return 1


def attach_peers(map_, rank, world, all_gather_object, deferred=False):
    """BGKOctoMap, one process per GPU: exchange the CUDA IPC handles of every replica's pool / flag buffer and attach
    them, so that the predict kernel of each rank stores its results straight into all replicas (include/la3dm_b200.h,
    "multi-GPU, BGKOctoMap").  all_gather_object(obj) -> list of every rank's obj, in rank order (e.g. a closure over
    torch.distributed.all_gather_object).  Call after map_.reserve_blocks()."""
    if world == 1:
        return
    if deferred:
        map_.peer_set_deferred(True)
    handles = all_gather_object(map_.peer_ipc_export())
    assert len(handles) == world
    pools, flags = [0] * world, [0] * world
    for p, (hp, hf) in enumerate(handles):
        if p != rank:
            pools[p], flags[p] = map_.peer_ipc_open(hp, hf)
    map_.peer_attach(world, rank, pools, flags)
