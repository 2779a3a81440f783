"""World-size-2 tests on CPU (gloo): the host-side logic of both multi-GPU modes.

No CUDA here: what is exercised is (a) the molecule-stream sharding bench.py uses for config 4 (disjoint, complete,
independent of the world size), (b) the row-slice formula of the sharded large-system path -- the Python mirror against
the library's own epnn_shard_slice, equal slices that tile [0, n) -- and an emulation of the in-place all-gather of equal
slices the library performs with NCCL (every rank ends with every owner's rows, bit for bit), and (c) the hand-over of
rank 0's NCCL unique id to every rank through torch.distributed (gloo here, nccl on the GPUs)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from epnn_b200 import shard, synth
        # (a) molecule stream: rank r generates molecules [r*M, (r+1)*M) of the same stream
        M = 300
        offs, xyz, sp, Q = synth.qm9_shaped(M, 9, seed=0, first=rank * M)
        full = synth.qm9_shaped(world * M, 9, seed=0, first=0)
        a0, a1 = full[0][rank * M], full[0][(rank + 1) * M]
        assert np.array_equal(xyz, full[1][a0:a1]) and np.array_equal(sp, full[2][a0:a1])
        assert np.array_equal(offs, full[0][rank * M:(rank + 1) * M + 1] - a0)
        counts = torch.tensor([int(offs[-1])], dtype=torch.int64)
        dist.all_reduce(counts)
        assert int(counts.item()) == int(full[0][-1])
        # (b) row slices: equal sizes (multiples of 32), tile [0, n), agree with the library's own formula
        from epnn_b200 import _capi
        lib = _capi.load()
        for n in (1, 7, 2220, 2220 + 700 + 90, 1_000_000, 2 ** 31 - 1):
            b, e = shard.slice_rows(n, rank, world)
            cb, ce = C.c_int64(-1), C.c_int64(-1)
            assert lib.epnn_shard_slice(n, rank, world, C.byref(cb), C.byref(ce)) == 0
            assert (cb.value, ce.value) == (b, e)
            lo = torch.tensor([b, e], dtype=torch.int64)
            allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(allr, lo)
            assert allr[0][0] == 0 and allr[-1][1] == n
            assert all(int(allr[k][1]) == int(allr[k + 1][0]) for k in range(world - 1))
            size = -(-n // world); size = (size + 31) // 32 * 32
            assert all(int(a[1] - a[0]) in (size, max(0, n - size * k)) for k, a in enumerate(allr))
        assert lib.epnn_shard_slice(10, 2, 2, C.byref(cb), C.byref(ce)) == -1        # rank out of range
        # in-place all-gather of equal slices, the exchange the library issues after every step / pass: each rank fills the
        # rows it owns of a padded array, the gather must hand every rank the complete array
        rng = np.random.default_rng(42)
        n = 2220 + 700 + 90
        ref = rng.normal(size=(n, 32)).astype(np.float32)               # same on both ranks (same seed)
        size = (-(-n // world) + 31) // 32 * 32
        mine = np.full((size * world, 32), np.nan, np.float32)
        b, e = shard.slice_rows(n, rank, world)
        mine[b:e] = ref[b:e]
        parts = [torch.zeros(size, 32) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine[size * rank:size * (rank + 1)]))
        full = torch.cat(parts).numpy()
        assert np.array_equal(full[:n], ref)
        # (c) the NCCL unique id travels from rank 0 to every rank (gloo: host tensor).  ncclGetUniqueId needs no GPU; if this
        # box has no usable libnccl the library says so instead of handing out garbage
        got = None
        try:
            got = shard.broadcast_unique_id(lib, rank)
        except _capi.EpnnError as ex:
            assert rank == 0 and "nccl" in str(ex).lower()
        flag = torch.tensor([0 if got is None else 1])
        dist.all_reduce(flag)
        if int(flag.item()) == world:
            ids = [torch.zeros(128, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(ids, torch.frombuffer(bytearray(bytes(got)), dtype=torch.uint8))
            assert all(torch.equal(ids[0], x) for x in ids) and int(ids[0].sum()) > 0
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_init_argument_checks():
    """No GPU needed: a NULL ctx is rejected before anything else."""
    from epnn_b200 import _capi
    lib = _capi.load()
    assert lib.epnn_shard_init(None, 0, 2, None) == -1
    assert lib.epnn_shard_unique_id(None) == -1
