"""Multi-GPU plumbing (host side): work slicing and the all-reduce callback that ``epnn_set_shard`` calls.

Two ways to use several GPUs:

* many independent systems (BASELINE config 4): every rank runs its own contiguous range of the batch through its
  own context; no collective (``molecule_range``).
* one big system (config 5): every rank gets the same ``infer_batch`` call; the library splits the large-system
  pair kernels by ``slice_range`` and, once per message-passing step / electron-passing pass, calls back into
  :func:`make_allreduce` -- ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs, gloo in the CPU tests) on
  the partial-sum / charge-transfer buffer.  Each element is non-zero on exactly one rank, so the result is exact.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


def slice_range(n: int, rank: int, world: int):
    """[begin, end) of rank's contiguous share of n work units: the formula libepnn_b200 uses (64-bit floor division)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def molecule_range(n_systems: int, rank: int, world: int):
    return slice_range(n_systems, rank, world)


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view a raw device pointer without copying."""

    def __init__(self, ptr, count, is_double):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8" if is_double else "<f4",
                                         "data": (int(ptr), False), "version": 2, "strides": None}


def make_allreduce(group=None, device=None, stream_handle: int = 0):
    """Returns ``(ctypes callback, state)``.  ``device`` None = host pointers (gloo tests); otherwise a CUDA device
    index and the ctx stream handle: the collective is enqueued on that stream, after the kernels that produced the
    buffer and before the ones that consume it."""
    import torch
    import torch.distributed as dist
    state = {"calls": 0, "bytes": 0, "error": None}
    ext = None
    dev = None
    if device is not None:
        dev = torch.device("cuda", device)
        ext = torch.cuda.ExternalStream(stream_handle, device=dev)

    def _cb(user, ptr, count, is_double, stream):
        try:
            if dev is None:
                ct = C.c_double if is_double else C.c_float
                arr = np.ctypeslib.as_array((ct * count).from_address(ptr))
                dist.all_reduce(torch.from_numpy(arr), group=group)
            else:
                t = torch.as_tensor(_DevArray(ptr, count, is_double), device=dev)
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, group=group)
            state["calls"] += 1
            state["bytes"] += int(count) * (8 if is_double else 4)
            return 0
        except Exception as ex:      # noqa: BLE001 -- never raise across the C boundary
            state["error"] = ex
            return 1

    return _capi.ALLREDUCE_FN(_cb), state
