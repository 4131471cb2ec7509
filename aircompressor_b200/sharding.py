"""Per-GPU batch split.  Blocks are independent (AbstractTestCompression.java:376-382 proves the
codecs carry no state between calls), so multi-GPU is a contiguous partition of the block index range,
balanced by bytes, with no collective on the data path."""
import numpy as np


def partition_by_bytes(sizes, world_size):
    """Splits blocks [0, n) into `world_size` contiguous ranges whose byte totals are as even as a
    contiguous split allows.  Returns a list of (begin, end) index pairs covering [0, n) in order."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = sizes.size
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        cut = int(np.searchsorted(csum, target, side="left"))
        cut = max(bounds[-1], min(n, cut))
        bounds.append(cut)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def shard_for_rank(sizes, rank, world_size):
    return partition_by_bytes(sizes, world_size)[rank]
