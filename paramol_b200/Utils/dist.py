# -*- coding: utf-8 -*-
"""
Description
-----------
Minimal ``torch.distributed`` plumbing for the conformation-sharded objective: one process per GPU, every rank
holds a contiguous shard of each system's conformations and the ONLY data-path collective is a sum-allreduce of a
small float64 buffer.  When no process group is initialised every helper is the identity,
so single-GPU code pays nothing.  Works with the ``nccl`` backend (CUDA tensors, NVLink) and with ``gloo``
(CPU tensors; used by the world_size-2 CPU tests).
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def backend():
    d = _dist()
    return d.get_backend() if d else None


def allreduce_sum_numpy(array):
    """Sum a float64 numpy array over ranks (setup-time reductions: variances, covariances, counts)."""
    d = _dist()
    a = np.ascontiguousarray(array, dtype=np.float64)
    if d is None or d.get_world_size() == 1:
        return a
    import torch
    t = torch.from_numpy(a.copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_sum_tensor_(tensor):
    """In-place sum-allreduce of a torch tensor (the per-call partial-sum buffer); identity on one rank."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return tensor
    if d.get_backend() == "gloo" and tensor.is_cuda:
        t = tensor.cpu()
        d.all_reduce(t, op=d.ReduceOp.SUM)
        tensor.copy_(t)
        return tensor
    d.all_reduce(tensor, op=d.ReduceOp.SUM)
    return tensor


def allreduce_min_tensor_(tensor):
    """In-place min-allreduce of a torch tensor; identity on one rank."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return tensor
    if d.get_backend() == "gloo" and tensor.is_cuda:
        t = tensor.cpu()
        d.all_reduce(t, op=d.ReduceOp.MIN)
        tensor.copy_(t)
        return tensor
    d.all_reduce(tensor, op=d.ReduceOp.MIN)
    return tensor


def shard_bounds(n_items, n_ranks=None, r=None):
    """Contiguous shard [lo, hi) of rank ``r``: the first ``n_items % n_ranks`` ranks get one extra item."""
    n_ranks = world_size() if n_ranks is None else n_ranks
    r = rank() if r is None else r
    base, extra = divmod(int(n_items), int(n_ranks))
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)
