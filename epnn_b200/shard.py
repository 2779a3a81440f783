"""Multi-GPU plumbing (host side).

Two ways to use several GPUs:

* many independent systems (BASELINE config 4): every rank runs its own contiguous range of the batch through its
  own context; no collective (``molecule_range``).
* one big system (config 5): every rank gets the same ``infer_batch`` call after ``Engine.set_shard(rank, world)``.
  The library cuts the atom index space into equal slices (``slice_rows`` mirrors ``epnn_shard_slice``), each rank
  owns the rows of its slice, and the per-step / per-pass exchanges are ``ncclAllGather`` calls made by the library
  itself on the ctx stream (NCCL is dlopen-ed there).  The only thing the host has to do is hand rank 0's NCCL unique
  id to every rank: ``broadcast_unique_id`` does it with ``torch.distributed`` -- nccl or gloo, it is 128 bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def molecule_range(n_systems: int, rank: int, world: int):
    """[begin, end) of rank's contiguous share of n independent systems (64-bit floor division)."""
    return (n_systems * rank) // world, (n_systems * (rank + 1)) // world


def slice_rows(n_atoms: int, rank: int, world: int):
    """Rows [begin, end) of a batch of ``n_atoms`` atoms that ``rank`` owns in a sharded call: equal slices of
    ceil(n / world) rows rounded up to 32, clipped to n.  Same formula as ``epnn_shard_slice`` (tested against it)."""
    s = -(-n_atoms // world)
    s = (s + 31) // 32 * 32
    return min(n_atoms, s * rank), min(n_atoms, s * (rank + 1))


def broadcast_unique_id(lib, rank: int, group=None, device=None):
    """Rank 0 asks the library for a fresh ncclUniqueId (``epnn_shard_unique_id``); every rank returns the same 128 bytes
    as a ctypes buffer.  ``device``: CUDA device index when the process group's backend is nccl (the broadcast tensor
    must then live on the GPU), None for gloo."""
    import torch
    import torch.distributed as dist
    buf = (C.c_char * 128)()
    if rank == 0:
        rc = lib.epnn_shard_unique_id(buf)
        if rc != 0:
            from . import _capi
            raise _capi.EpnnError(rc, lib.epnn_last_error(None).decode())
    t = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy())
    on_gpu = device is not None and dist.get_backend(group) == "nccl"
    if on_gpu:
        t = t.to(torch.device("cuda", device))
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    out = (C.c_char * 128).from_buffer_copy(t.cpu().numpy().tobytes())
    return out
