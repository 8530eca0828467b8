"""Multi-GPU plumbing for independent trials: one process per GPU, contiguous ranges of the global
trial index per rank, no data-path collective.  The only exchanges are (a) a host gather of the
per-rank results and (b) the sum of the per-noise-level error accumulators of experiments.m:112-124,
done as gather + addition in rank order on the root so the means are bit-stable for a given world size.
``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) is used only for that plumbing."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of `total` items for `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def gather_arrays(local, dst=0):
    """Gather a dict of NumPy arrays (concatenated along axis 0 in rank order) on `dst`; None elsewhere."""
    d = _dist()
    if d is None:
        return dict(local)
    rank, size = d.get_rank(), d.get_world_size()
    bucket = [None] * size if rank == dst else None
    d.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    return {k: np.concatenate([b[k] for b in bucket], axis=0) for k in local}


def sum_in_rank_order(partial, dst=0):
    """Sum equally-shaped NumPy accumulators over ranks, adding in rank order on `dst` (deterministic)."""
    d = _dist()
    if d is None:
        return np.array(partial, copy=True)
    rank, size = d.get_rank(), d.get_world_size()
    bucket = [None] * size if rank == dst else None
    d.gather_object(np.asarray(partial), bucket, dst=dst)
    if rank != dst:
        return None
    total = np.array(bucket[0], copy=True)
    for b in bucket[1:]:
        total = total + b
    return total


def max_over_ranks(value):
    """Max of a Python float over ranks (device timings are reported as the slowest rank's)."""
    d = _dist()
    if d is None:
        return float(value)
    import torch
    dev = "cuda" if d.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())
