"""Multi-GPU plumbing of self-play: games shard across ranks, nothing is exchanged on the data path.

The reference fans its game-index list out to worker processes in contiguous slices (selfplay_main.py:44-47); here a
worker is one process per GPU (torchrun / torch.distributed), each with its own board pool, node pool and weight
replica.  The only collective is the reduction of the throughput counters at the end of a run.
"""
import math
import os


def split_indices(num_data, parts):
    """selfplay_main.py:44-47: indices 1..num_data in contiguous slices of ceil(num_data / parts)."""
    index_list = list(range(1, num_data + 1))
    size = math.ceil(num_data / parts) if parts > 0 else num_data
    return [index_list[i:i + size] for i in range(0, len(index_list), size)]


def shard_for_rank(num_data, rank=None, world=None):
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    parts = split_indices(num_data, world)
    return parts[rank] if rank < len(parts) else []


def reduce_counters(moves, seconds, device=None):
    """Whole-job moves and the slowest rank's time (max over ranks): moves/sec = sum(moves) / max(seconds)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(moves), float(seconds)
    t = torch.tensor([float(moves), float(seconds)], dtype=torch.float64, device=device or "cpu")
    s = t.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
    m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s[0].item(), m[1].item()
