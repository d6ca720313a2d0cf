"""Multi-GPU sharding of independent planning queries: the primitive library is replicated on every
GPU, the query batch is split into contiguous blocks, one process per GPU decides its block, and only the result
bits travel (gather to rank 0).  There is no collective on the data path."""
import numpy as np


def block_bounds(n, world):
    """contiguous partition of n items over `world` ranks: rank r gets [bounds[r], bounds[r+1])"""
    base, rem = divmod(n, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_queries(queries, rank, world):
    b = block_bounds(len(queries), world)
    return queries[b[rank]:b[rank + 1]]


def run_sharded(backend, queries, cand_idx=None, mode=0, group=None):
    """every rank calls this with the FULL query array; returns the full collision-bit array on rank 0 (None elsewhere).
    `backend` is this rank's CollisionEngine (library and scenes already staged on it)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        bits, st = backend.query(queries, cand_idx, mode)
        return bits, [st]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_queries(queries, rank, world)
    bits, st = backend.query(mine, cand_idx, mode)
    packed = np.packbits(bits.astype(np.uint8), bitorder='little')
    out = [None] * world if rank == 0 else None
    dist.gather_object((packed, len(bits), st), out, dst=0, group=group)      # result bitmasks only
    if rank != 0:
        return None, None
    parts = [np.unpackbits(p, bitorder='little')[:n] for p, n, _ in out]
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8), [s for _, _, s in out]
