"""Host-side helpers for the multi-GPU path: point-block sharding + NCCL bootstrap.

One process per GPU.  The source cloud is split into contiguous blocks; every rank reduces its block, the
27 + 5 accumulators are summed with ONE 32-double all-reduce on the context's stream (inside the C library,
`dcreg_comm_init`), and every rank runs the solve redundantly on identical inputs.  `torch.distributed` is only the
plumbing that carries the 128-byte NCCL unique id from rank 0 to the others.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n` source slots owned by `rank`; blocks differ by at most one slot and
    tile [0, n) exactly."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    lo = rank * n // world
    hi = (rank + 1) * n // world
    return lo, hi


def broadcast_unique_id(make_id, dist, device=None) -> bytes:
    """Rank 0 calls `make_id()` (-> 128 bytes, `Context.comm_unique_id`); everyone returns the same bytes."""
    import torch
    rank = dist.get_rank()
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf = torch.tensor(list(raw), dtype=torch.uint8, device=buf.device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def init_sharded(ctx, dist, n_total: int, device=None):
    """Attach an NCCL communicator to `ctx` for the current torch.distributed world and record the global slot count."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world > 1:
        uid = broadcast_unique_id(ctx.comm_unique_id, dist, device)
        ctx.comm_init(uid, rank, world)
    ctx.set_global_source_count(n_total)
    return shard_range(n_total, rank, world)
