"""Channel sharding for multi-GPU runs.

Channels are fully independent (every reference demod instance owns all of its state,
/root/reference/src/gr/gr_demod_base.h:144-184), so the path shards by channel index with NO data-path
collective: rank r of W demodulates its own block of channels on its own GPU.  A mixed channel list
(BASELINE config 4: FM / 4FSK / QPSK interleaved) is first grouped by mode -- one qrl_rx handle per (rank, mode),
so a warp never mixes modes -- then block-partitioned across ranks mode by mode.

`gather_counts` is the only collective a sharded receiver needs for bookkeeping (per-channel produced-item
counts back on rank 0); it works over any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from collections import OrderedDict


def partition(n_items, world, rank):
    """Contiguous block partition: items [lo, hi) for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_channels(modes, world, rank):
    """modes: list with one hashable mode key per global channel (e.g. ("4fsk", 5, 3000, True)).
    Returns OrderedDict mode -> list of global channel indices this rank owns (sorted, contiguous per mode)."""
    by_mode = OrderedDict()
    for ch, m in enumerate(modes):
        by_mode.setdefault(m, []).append(ch)
    out = OrderedDict()
    for m, chans in by_mode.items():
        lo, hi = partition(len(chans), world, rank)
        if hi > lo:
            out[m] = chans[lo:hi]
    return out


def owner_of(modes, world):
    """global channel index -> owning rank (inverse of shard_channels)."""
    owner = [None] * len(modes)
    for r in range(world):
        for chans in shard_channels(modes, world, r).values():
            for ch in chans:
                owner[ch] = r
    return owner


def gather_counts(local_counts, modes, world, rank, dist, device="cpu"):
    """All ranks contribute {global channel: produced items}; returns the full list on every rank."""
    import torch
    t = torch.zeros(len(modes), dtype=torch.int64, device=device)
    for ch, n in local_counts.items():
        t[ch] = int(n)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()
