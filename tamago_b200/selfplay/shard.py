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


def shard_offset(rank=None, span=10_000_000):
    """First game index (minus one) of a rank when ranks take disjoint, contiguous index ranges of `span` games each:
    open-ended runs (benchmarks, continuous self-play) where num_data is not known in advance."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    return rank * span


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


def gather_training_arrays(arrays, dst=0):
    """Optional collection of finished training records on one rank (the reference's workers meet in a shared
    directory instead): `arrays` is a dict of numpy arrays whose first dimension is the local sample count.  Uses one
    all_gather of the counts and one padded all_gather per array (NCCL when the tensors live on a GPU, gloo on the CPU)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return arrays
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = len(next(iter(arrays.values())))
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    nmax = max(counts + [1])
    out = {}
    for key, a in arrays.items():
        a = np.ascontiguousarray(a)
        pad = np.zeros((nmax,) + a.shape[1:], a.dtype)
        pad[:n_local] = a
        t = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        if rank == dst:
            out[key] = np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)])
    return out if rank == dst else None


def gather_sample_tensors(tensors):
    """north_star: "an optional NCCL gather collects finished npz records".  `tensors`: the device-resident sample arrays of
    this rank (Engine.sample_tensors(): input f32 [n, 6, N, N], policy f64 [n, A], value i32 [n]); returns the same arrays
    for ALL ranks' samples, in rank order, on every rank -- one all_gather of the counts and one padded all_gather per
    array over NCCL (NVLink / NVSwitch), nothing touches the host."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuple(t.clone() for t in tensors), [len(tensors[0])]
    world = dist.get_world_size()
    dev = tensors[0].device
    n_local = torch.tensor([len(tensors[0])], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts + [1])
    out = []
    for t in tensors:
        pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:len(t)] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out.append(torch.cat([p[:c] for p, c in zip(parts, counts)]))
    return tuple(out), counts
