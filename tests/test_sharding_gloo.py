"""Host-side logic of the multi-GPU exchange (la3dm_b200/sharding.py) with world_size 2 over gloo on CPU: test blocks
dealt round-robin, fixed-size rows, ONE all_gather per scan, peers' rows scattered back.  The pack / unpack kernels are
replaced by a numpy stand-in with the same indexing (la3dm_b200/csrc/shard.cu); the GPU path itself is covered by
tests/test_gpu_bgk.py::test_two_rank_sharding_single_gpu."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

ROW = 48


class FakeMap:
    """Replica of a map as an array [T, ROW] of block records; rank r 'predicts' (modifies) rows t % world == r."""

    def __init__(self, rank, world, T, seed):
        self.rank, self.world, self.T = rank, world, T
        rng = np.random.default_rng(seed)
        self.blocks = rng.integers(0, 255, (T, ROW), dtype=np.uint8)      # identical on every rank
        mine = np.arange(T) % world == rank
        self.blocks[mine] = (self.blocks[mine].astype(np.int32) + 1 + rank).astype(np.uint8)   # this rank's update
        self.bufs = {}

    def shard_rows(self):
        return (self.T + self.world - 1) // self.world, ROW

    def alloc(self, nbytes):
        t = torch.zeros(nbytes, dtype=torch.uint8)
        self.bufs[t.data_ptr()] = t
        return t, t.data_ptr()

    def shard_pack(self, ptr):
        rows, _ = self.shard_rows()
        buf = self.bufs[ptr].numpy().reshape(rows, ROW)
        for r in range(rows):
            t = r * self.world + self.rank
            if t < self.T:
                buf[r] = self.blocks[t]

    def shard_unpack(self, ptr):
        rows, _ = self.shard_rows()
        buf = self.bufs[ptr].numpy().reshape(self.world, rows, ROW)
        for q in range(self.world):
            if q == self.rank:
                continue
            for r in range(rows):
                t = r * self.world + q
                if t < self.T:
                    self.blocks[t] = buf[q, r]


def _worker(rank, world, port, T, out):
    sys.path.insert(0, ROOT)
    from la3dm_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = FakeMap(rank, world, T, seed=7)
    n = sharding.exchange(m, world, lambda o, i: dist.all_gather_into_tensor(o, i), m.alloc)
    assert n == (1 if T else 0)
    # expected: every row carries its owner's update
    rng = np.random.default_rng(7)
    want = rng.integers(0, 255, (T, ROW), dtype=np.uint8)
    for t in range(T):
        want[t] = (want[t].astype(np.int32) + 1 + t % world).astype(np.uint8)
    ok = np.array_equal(m.blocks, want)
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        out.put(all(gathered))
    dist.destroy_process_group()


def _run(T, port):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_index_maps():
    from la3dm_b200 import sharding
    for world in (1, 2, 4, 8):
        for T in (0, 1, 7, 64, 1001):
            rows = sharding.rows_per_rank(T, world)
            assert rows * world >= T and (rows - 1) * world < max(T, 1)
            seen = set()
            for t in range(T):
                r, q = sharding.row_of(t, world), sharding.owner_of(t, world)
                assert sharding.test_block_of(q, r, world) == t and r < rows
                seen.add((q, r))
            assert len(seen) == T


def test_exchange_world2_gloo_odd_and_empty():
    _run(1001, 29631)     # ragged: the last row of rank 1 is padding
    _run(0, 29632)        # a scan without test blocks issues no collective


class FakePeerMap:
    """Stands in for la3dm_b200.BGKOctoMap in attach_peers(): handles are (rank-tagged) byte strings, 'opening' one
    returns pointers derived from the tag, so the test can check who attached what."""

    def __init__(self, rank):
        self.rank, self.attached = rank, None

    def peer_ipc_export(self):
        return (b"P%03d" % self.rank).ljust(64, b"\0"), (b"F%03d" % self.rank).ljust(64, b"\0")

    def peer_ipc_open(self, hp, hf):
        assert hp[:1] == b"P" and hf[:1] == b"F" and hp[1:4] == hf[1:4]
        q = int(hp[1:4])
        assert q != self.rank                       # a rank never opens its own handle (cudaIpcOpenMemHandle would fail)
        return 0x1000 + q, 0x2000 + q

    def peer_attach(self, world, rank, pools, flags):
        self.attached = (world, rank, list(pools), list(flags))


def _peer_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from la3dm_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(obj):
        lst = [None] * world
        dist.all_gather_object(lst, obj)
        return lst

    m = FakePeerMap(rank)
    sharding.attach_peers(m, rank, world, gather)
    w, r, pools, flags = m.attached
    ok = w == world and r == rank and all((pools[q], flags[q]) == ((0, 0) if q == rank else (0x1000 + q, 0x2000 + q))
                                          for q in range(world))
    res = gather(bool(ok))
    if rank == 0:
        out.put(all(res))
    dist.destroy_process_group()


def test_attach_peers_world2_gloo():
    """la3dm_b200.sharding.attach_peers: every rank exports its two IPC handles, ONE all_gather_object moves them, every
    rank opens the others' (never its own) and attaches pointer tables indexed by rank."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, 29633, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
