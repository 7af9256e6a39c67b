"""Batch sharding for N GPUs of one box: images are independent units, so each rank takes a
contiguous slice and there is no data-path collective.  torch.distributed is used for barriers / max-over-ranks timing."""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous, balanced slice [lo, hi) of `total` images for `rank` (first `total % world` ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world {world}')
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device='cpu'):
    """The slowest rank's time (ms): what a multi-GPU throughput number must be divided by."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, ms_local, device='cpu'):
    """Whole-job units/s = sum over ranks of units / max over ranks of time."""
    n = torch.tensor([float(units_per_rank)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()) / (max_over_ranks(ms_local, device) * 1e-3)
