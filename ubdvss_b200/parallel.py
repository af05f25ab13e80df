"""Multi-GPU plumbing (one process per GPU, torch.distributed): batch sharding for inference (no
collective -- images are independent, model_runner.py:127-134) and the gradient all-reduce of
data-parallel training (the only exchange step on the path; the reference itself is single-device)."""
from __future__ import annotations


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) slice of a batch for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(flat, dist=None, group=None) -> float:
    """In-place SUM all-reduce of the flat gradient tensor; returns the 1/world factor that Adam
    applies (`ubd_adam_step(..., grad_scale)`), so the averaged gradient is never materialised."""
    if dist is None:
        import torch.distributed as dist
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)
