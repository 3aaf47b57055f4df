"""Host-side logic of the batch-dimension data parallelism: who owns which rows / lanes.

The reference's training step sums per-example gradients (gad/src/net_ext.rs:62-65), so the rows of a batch can
be split over ranks and the weight gradients (and the loss) summed afterwards: the result differs from the
single-process one only by the association of that sum. Nothing here touches a device, so it is covered by
world-size-2 `gloo` tests on CPU (tests/test_dp_gloo.py) with the oracle standing in for the kernels.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of `rank`'s contiguous share of `total` units; the first `total % world` ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(total: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(total, world, r) for r in range(world)]


def even_rows(total: int, world: int) -> int:
    """Rows per rank when the batch must split evenly (the tensor-core GEMM wants equal, 256-aligned shards)."""
    if total % world:
        raise ValueError(f"batch of {total} rows does not split evenly over {world} ranks")
    return total // world
