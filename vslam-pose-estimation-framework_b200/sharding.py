"""Multi-GPU layout of the hot path: independent units only, no data-path collective.

Within one sequence, frame t+1 depends on frame t through the FAST threshold controller and the motion prior
(/root/reference/src/framepoint_generation/base_framepoint_generator.cpp:440-459,
/root/reference/src/position_tracking/pose_tracker_3d.cpp:46,239), so a sequence never spans GPUs ("replicas only").
Independent stereo pairs / sequences shard one block per rank; results go back to the host of their own rank.
torch.distributed is used only for the barrier and the max-over-ranks of the timings (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations


def block_partition(n_units: int, world: int, rank: int) -> range:
    """contiguous block partition: unit i -> rank floor(i * world / n_units)"""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    lo = -(-rank * n_units // world)          # ceil(rank * n / world): first i with floor(i*world/n) == rank
    hi = -(-(rank + 1) * n_units // world)
    return range(lo, hi)


def sequence_owner(sequence: int, world: int) -> int:
    """independent sequences: sequence s -> rank s mod world"""
    return sequence % world


def weak_seeds(distinct_per_rank: int, rank: int) -> range:
    """weak scaling: every rank generates its own disjoint block of synthetic pair seeds"""
    return range(rank * distinct_per_rank, (rank + 1) * distinct_per_rank)


def max_over_ranks(value: float, device=None) -> float:
    """the job's time is the slowest rank's time"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
