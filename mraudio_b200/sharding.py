"""Data-parallel partitioning of the path's independent units (videos / queries) across ranks.

The reference shards with ``DistributedSampler`` (utils/trainer.py:74-75); every (video, frame) row of the Q-Former and
every query of the scorer is independent, so ranks take contiguous blocks and the forward needs no collective."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank``; block sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
