"""World-size-2 tests on CPU (gloo): the host-side logic of both multi-GPU modes.

No CUDA here: what is exercised is (a) the molecule-stream sharding bench.py uses for config 4 (disjoint, complete,
independent of the world size), (b) the slicing formula shared with libepnn_b200 and (c) the all-reduce callback
epnn_set_shard calls -- through the very ctypes function object the library would call -- with each rank owning a
slice of an otherwise zero buffer, which must reproduce the unsharded buffer bit for bit."""
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
        # (b) slicing covers [0, n) without gaps or overlap for awkward n
        for n in (0, 1, 7, 555 * 18, 10 ** 10 + 3):
            b, e = shard.slice_range(n, rank, world)
            lo = torch.tensor([b, e], dtype=torch.int64)
            allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(allr, lo)
            assert allr[0][0] == 0 and allr[-1][1] == n
            assert all(int(allr[k][1]) == int(allr[k + 1][0]) for k in range(world - 1))
        # (c) the callback, called the way the C library calls it
        cb, state = shard.make_allreduce(group=None, device=None)
        rng = np.random.default_rng(42)
        for dtype, is_double in ((np.float32, 0), (np.float64, 1)):
            ref = rng.normal(size=10007).astype(dtype)              # same on both ranks (same seed)
            ref[::13] = 0.0
            mine = np.zeros_like(ref)
            b, e = shard.slice_range(len(ref), rank, world)
            mine[b:e] = ref[b:e]
            rc = cb(None, mine.ctypes.data_as(C.c_void_p), len(mine), is_double, None)
            assert rc == 0 and state["error"] is None
            assert np.array_equal(mine, ref)                        # x + 0 is exact: bit-identical to the unsharded buffer
        assert state["calls"] == 2
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_set_shard_argument_checks():
    """No GPU needed: a NULL ctx is rejected before anything else."""
    from epnn_b200 import _capi
    lib = _capi.load()
    assert lib.epnn_set_shard(None, 0, 2, _capi.ALLREDUCE_FN(0), None) == -1
