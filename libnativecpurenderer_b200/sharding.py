"""Frame sharding across ranks: frames are independent when each starts by overwriting the
canvas, so N GPUs render disjoint frame sets with no data-path collective.  The only communication is the
measurement plumbing below (barrier + a max/sum of scalars), over ``torch.distributed`` — NCCL on the GPU box,
gloo in the CPU tests."""
from __future__ import annotations


def frames_for_rank(n_frames: int, rank: int, world: int, mode: str = "interleave") -> range | list[int]:
    """Frame indices rendered by ``rank``: ``f % world == rank`` (load balance) or contiguous blocks (GOP locality)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    if mode == "interleave":
        return range(rank, n_frames, world)
    if mode == "block":
        per, extra = divmod(n_frames, world)
        start = rank * per + min(rank, extra)
        return range(start, start + per + (1 if rank < extra else 0))
    raise ValueError(mode)


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, dist=None, device: str = "cpu") -> float:
    """Whole-job throughput: units of all ranks / slowest rank's time (the contract's max-over-ranks timing)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return units_this_rank / seconds_this_rank
    import torch

    t = torch.tensor([seconds_this_rank], dtype=torch.float64, device=device)
    u = torch.tensor([units_this_rank], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item() / t.item())
