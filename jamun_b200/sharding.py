"""Chain sharding for multi-GPU sampling: chains are independent, so each rank owns a contiguous block
and nothing is exchanged until the final gather of samples."""
from __future__ import annotations


def shard_chains(num_chains: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced block of chain indices for `rank` (first `num_chains % world_size` ranks get one more)."""
    base, rem = divmod(num_chains, world_size)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))
