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
