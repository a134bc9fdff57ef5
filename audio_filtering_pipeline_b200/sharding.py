"""Stream sharding across GPUs: streams are independent, so rank r of W owns a
contiguous range of stream ids and no collective touches the data path.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) is used for the barrier and for the max-over-ranks time only."""
from __future__ import annotations


def partition_streams(total_streams, world_size, rank, align=2):
    """Contiguous ``[first, first+count)`` owned by ``rank``.  Counts differ by at most ``align`` and
    every rank but the last gets a multiple of ``align`` (stereo pairs / kernel row pairs stay on one
    GPU).  Ranks past the end get count 0."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    units = (total_streams + align - 1) // align
    base, extra = divmod(units, world_size)
    first_u = rank * base + min(rank, extra)
    count_u = base + (1 if rank < extra else 0)
    first = min(first_u * align, total_streams)
    last = min((first_u + count_u) * align, total_streams)
    return first, last - first


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
