"""Batch sharding over independent replicas (one process per GPU).

Images are independent (inference-mode batch norm, per-image decode), so the
N-GPU path has no data-path collective: every rank runs the whole path on a
contiguous slice of the batch and rank 0 gathers the small per-image result
tables in image order.  ``torch.distributed`` is used only for
that gather (and, in bench.py, for the barrier and the max-reduction of the
elapsed time); the backend is NCCL on GPUs and gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice ``[start, stop)`` of ``n_items`` for ``rank``: the first ``n_items % world`` ranks get one
    extra item, so the slices differ by at most one image and concatenate back in order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_in_order(local: Sequence, dst: int = 0, group=None) -> Optional[List]:
    """Gathers every rank's list of per-image results on ``dst`` and concatenates them in rank (= image) order.
    Returns the full list on ``dst``, ``None`` elsewhere.  Works without an initialised process group
    (single replica)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out: List = []
    for part in bucket:
        out.extend(part)
    return out


def run_sharded(items: Sequence, infer: Callable[[Sequence], Sequence], dst: int = 0, group=None) -> Optional[List]:
    """Runs ``infer`` (a callable mapping a batch of items to one result per item, e.g. ``Engine.infer_host``)
    on this rank's slice of ``items`` and gathers all results on ``dst`` in the original order."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(len(items), world, rank)
    local = list(infer(items[lo:hi])) if hi > lo else []
    if len(local) != hi - lo:
        raise RuntimeError(f"infer returned {len(local)} results for {hi - lo} items")
    return gather_in_order(local, dst, group)
