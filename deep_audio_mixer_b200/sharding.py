"""Multi-GPU plumbing: one process per GPU, songs/segments sharded with no data-path collective.

The reference is single-process.  Units (songs for mixing/evaluation, stem-chunk rows
for featurisation) are independent, so ranks never exchange audio; the only collectives are small
statistics all-reduces -- per-track (sum of loudness, count) for ``compute_mean_loudness``
(data/dataset.py:115-130) and (sum, sum of squares, count) of dB features -- over NCCL (gloo in the
CPU tests).
"""
import numpy as np
import torch


def assign_lpt(costs, world_size: int):
    """Longest-processing-time-first assignment of units to ranks: returns a list of
    index lists, one per rank, with near-equal total cost."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for r in range(world_size):
        out[r].sort()
    return out


def shard_range(n_units: int, rank: int, world_size: int):
    """Contiguous equal-cost shard [lo, hi) of n_units units."""
    base, rem = divmod(n_units, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist_ready():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def allreduce_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce when a process group exists; identity otherwise."""
    if _dist_ready() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return t


def allreduce_sum_count(sums, counts, device=None):
    """All-reduce a (sums, counts) pair of small float64 vectors; returns numpy arrays."""
    v = torch.as_tensor(np.concatenate([np.asarray(sums, dtype=np.float64), np.asarray(counts, dtype=np.float64)]))
    if _dist_ready() and torch.distributed.get_world_size() > 1:
        if torch.distributed.get_backend() == 'nccl':
            v = v.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
        allreduce_(v)
    v = v.cpu().numpy()
    n = len(v) // 2
    return v[:n], v[n:]


def feature_mean_std(stats: torch.Tensor):
    """stats = (sum, sum of squares, count) accumulated by K1 on this rank -> global (mean, std)
    after an all-reduce (parity unpinned, the reference never computes it)."""
    s = stats.clone()
    if _dist_ready() and torch.distributed.get_world_size() > 1 and torch.distributed.get_backend() != 'nccl':
        s = s.cpu()
    allreduce_(s)
    s = s.cpu().numpy()
    mean = s[0] / s[2]
    var = max(s[1] / s[2] - mean * mean, 0.0)
    return float(mean), float(np.sqrt(var))
