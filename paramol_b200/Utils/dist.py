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


_LOCAL_ONLY = 0


class local_only:
    """Context manager: inside it every helper of this module behaves as on a single rank even though a process group exists
    (used to evaluate an unsharded copy of a workload on one rank as the parity reference of a sharded run)."""

    def __enter__(self):
        global _LOCAL_ONLY
        _LOCAL_ONLY += 1
        return self

    def __exit__(self, *exc):
        global _LOCAL_ONLY
        _LOCAL_ONLY -= 1
        return False


def _dist():
    if _LOCAL_ONLY:
        return None
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


# ---- device communicator: the all-reduce inside pm_eval_reduce (csrc/pm_comm.cu) ---------------------------------------
_COMMS = {}
COMM_MAX_VEC = 1024


class DeviceComm:
    """``pm_comm`` handle of this rank: inboxes of all ranks of the node mapped through CUDA IPC (NVLink peer memory)."""

    def __init__(self, handle, lib, max_vec):
        self.handle, self._lib, self.max_vec = handle, lib, max_vec

    def timeouts(self):
        return int(self._lib.pm_comm_timeouts(self.handle))

    def allreduce_(self, tensor):
        """In-place sum of a small float64 CUDA tensor (<= 8 * max_vec doubles) over ranks: one kernel, no NCCL call."""
        import ctypes
        import torch
        from .. import _lib
        assert tensor.is_cuda and tensor.dtype == torch.float64 and tensor.is_contiguous() and tensor.numel() <= 8 * self.max_vec
        stream = ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(tensor.device.index))
        _lib.check(self._lib.pm_comm_allreduce(self.handle, tensor.data_ptr(), tensor.numel(), stream), "pm_comm_allreduce")
        return tensor


def device_comm(device_index):
    """
    The :obj:`DeviceComm` of this rank for ``device_index``, or None: single rank, a backend other than NCCL, IPC not
    available (containers that hide the peers' devices), or ``PARAMOL_B200_COMM=nccl``.  Created on first use; every rank must
    call this at the same point (it is a collective: the IPC handles are exchanged with an all_gather).  When ANY rank fails to
    map its peers, all ranks agree on None and the caller uses the NCCL all-reduce instead (logged, never silent per rank).
    """
    import os
    d = _dist()
    if d is None or d.get_world_size() == 1 or d.get_backend() != "nccl" or os.environ.get("PARAMOL_B200_COMM", "p2p") == "nccl":
        return None
    if device_index in _COMMS:
        return _COMMS[device_index]
    import ctypes
    import logging
    import torch
    from .. import _lib
    L = _lib.lib()
    world, me = d.get_world_size(), d.get_rank()
    dev = torch.device("cuda", device_index)
    raw = ctypes.create_string_buffer(64)
    h = ctypes.c_void_p()
    ok = L.pm_comm_create(device_index, me, world, COMM_MAX_VEC, raw, ctypes.byref(h)) == 0
    why = "" if ok else L.pm_last_error().decode()
    mine = torch.tensor(list(raw.raw) + [1 if ok else 0], dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    d.all_gather(gathered, mine)
    table = torch.stack(gathered).cpu().numpy()
    comm = None
    if table[:, 64].all():
        handles = table[:, :64].tobytes()
        ok = L.pm_comm_connect(h, handles) == 0
        if not ok:
            why = L.pm_last_error().decode()
    else:
        ok = False
    flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    d.all_reduce(flag, op=d.ReduceOp.MIN)
    if float(flag.item()) == 1.0:
        comm = DeviceComm(h, L, COMM_MAX_VEC)
    else:
        logging.warning("paramol_b200: peer-memory communicator unavailable (%s); using the NCCL all-reduce", why or "a peer failed")
        if h:
            L.pm_comm_destroy(h)
    _COMMS[device_index] = comm
    return comm


def comm_kind(device_index=None):
    """"single", "p2p" (one-shot all-reduce inside the reduction kernel over NVLink peer memory), "nccl" or "gloo"."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return "single"
    if d.get_backend() != "nccl":
        return d.get_backend()
    if device_index is None:
        import torch
        device_index = torch.cuda.current_device()
    return "p2p" if _COMMS.get(device_index) is not None else "nccl"


def shard_bounds(n_items, n_ranks=None, r=None):
    """Contiguous shard [lo, hi) of rank ``r``: the first ``n_items % n_ranks`` ranks get one extra item."""
    n_ranks = world_size() if n_ranks is None else n_ranks
    r = rank() if r is None else r
    base, extra = divmod(int(n_items), int(n_ranks))
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)
