"""Data-parallel replicas for inference: observations are independent, so a global batch is split
contiguously across ranks, every rank holds a full bf16 replica and there is NO collective on the data path.
(ZoeDepth's metric-head router votes over the *local* batch, so parity is defined per shard -- exactly what the
reference would compute if it were handed that shard.)"""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int):
    """Slices every per-observation tensor of a model-input dict; a shared (3,3) intrinsic is passed through."""
    n = batch["input_ids"].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    out = {}
    for k, v in batch.items():
        if hasattr(v, "shape") and v.dim() >= 1 and v.shape[0] == n and not (k == "intrinsic" and v.dim() == 2):
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out
