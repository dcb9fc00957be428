"""CPU tier: the N>1 path (channel sharding) on world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qradiolink_b200 import sharding


def test_partition_covers_everything():
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            spans = [sharding.partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_mixed_modes_grouped_then_partitioned():
    modes = [("nbfm", "4fsk", "qpsk")[ch % 3] for ch in range(1024)]          # BASELINE config 4
    owner = sharding.owner_of(modes, 8)
    assert None not in owner
    for r in range(8):
        mine = sharding.shard_channels(modes, 8, r)
        assert set(mine) == {"nbfm", "4fsk", "qpsk"}
        assert 126 <= sum(len(v) for v in mine.values()) <= 129
        for m, chans in mine.items():
            assert all(modes[c] == m for c in chans) and chans == sorted(chans)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    modes = [("4fsk", "qpsk")[ch % 2] for ch in range(10)]
    mine = sharding.shard_channels(modes, world, rank)
    local = {ch: 100 + ch for chans in mine.values() for ch in chans}      # stand-in for per-channel bit counts
    full = sharding.gather_counts(local, modes, world, rank, dist)
    # max-over-ranks of a per-rank time, as bench.py does
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, full, float(t.item()), sorted(local)))
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    owned = []
    for rank, full, tmax, mine in res:
        assert full == [100 + ch for ch in range(10)] and tmax == 2.0
        owned += mine
    assert sorted(owned) == list(range(10))
