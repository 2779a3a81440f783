"""CPU emulation of the neighbour-list and radial-descriptor kernels (epnn_b200/csrc/epnn_neighbor.cu): the unmodified
CUDA source, run thread by thread (tools/emu), must give the oracle's lists BIT FOR BIT -- CSR of the e != 0 set with
ascending columns, the unordered pair list with its float64 distances, the reverse-edge index, the reference's is_near
flags (charge_gn.py:90-94) -- on the direct path (n <= 512 atoms) and on the cell-list path (n > 512), and the
descriptors / their rank-16 coefficients to float32 round-off (charge_gn.py:122-163)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from epnn_b200 import _capi
from oracle import epnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_neighbor.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU",
                           "-Wno-unknown-pragmas", "-o", LIB, os.path.join(ROOT, "tools", "emu", "emu_neighbor.cpp")])
    lib = C.CDLL(LIB)
    lib.emu_neighbor_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 12 + [C.c_int]
    lib.emu_neighbor_fill.argtypes = [C.c_int] + [C.c_void_p] * 15 + [C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(emu, offs, xyz, Q, ek):
    n, n_sys = int(offs[-1]), len(offs) - 1
    lib = _capi.load()
    mu = np.zeros(48); B = np.zeros((48, 16))
    assert lib.epnn_rbf_centers(_p(mu)) == 0 and lib.epnn_rbf_basis(_p(B)) == 0
    i32 = lambda k: np.zeros(k, np.int32)
    atom_sys, deg, degU, rowptr, ustart = i32(n), i32(n), i32(n), i32(n + 1), i32(n + 1)
    q0 = np.zeros(n)
    cap = 4 * n + 64 * n_sys + 2
    grid, cell_start, cell_atoms = i32(8 * (n_sys + 1)), i32(cap + 2), i32(n + 1)
    assert emu.emu_neighbor_counts(_p(mu), _p(B), n, n_sys, _p(offs), _p(xyz), _p(Q), _p(atom_sys), _p(q0), _p(deg), _p(degU),
                                   _p(rowptr), _p(ustart), _p(grid), _p(cell_start), _p(cell_atoms), cap) == 0
    nnz, P = int(rowptr[n]), int(ustart[n])
    col, pid, pair_i, pair_j = i32(nnz + 1), i32(nnz + 1), i32(P + 1), i32(P + 1)
    pair_D, Dtmp = np.zeros(P + 1), np.zeros(nnz + 1)
    e = np.zeros((P + 1, ek), np.float32); near = np.zeros(P + 16, np.uint8)
    assert emu.emu_neighbor_fill(n, _p(atom_sys), _p(offs), _p(xyz), _p(degU), _p(rowptr), _p(ustart), _p(grid), _p(cell_start),
                                 _p(cell_atoms), _p(col), _p(pid), _p(pair_i), _p(pair_j), _p(pair_D), _p(Dtmp), ek, _p(e), _p(near)) == 0
    return dict(atom_sys=atom_sys, q0=q0, rowptr=rowptr, col=col[:nnz], pid=pid[:nnz], pair_i=pair_i[:P], pair_j=pair_j[:P],
                pair_D=pair_D[:P], e=e[:P], near=near[:P], B=B, P=P)


def _batch(mixed, protein):
    """Three molecules (direct search), a 100-atom protein cut (direct search, large system), a 700-atom cut (cell list)."""
    offs, xyz, _, Q = mixed.batch(mixed.usable(9)[[2, 900, 3500]].tolist(), 9)
    offs = np.concatenate([offs, [offs[-1] + 100, offs[-1] + 800]]).astype(np.int32)
    xyz = np.concatenate([xyz, protein["xyz"][1000:1100], protein["xyz"][:700]]).astype(np.float32)
    Q = np.concatenate([Q, [1.0, -2.0]]).astype(np.float32)
    return offs, xyz, Q


def test_emulated_neighbor_lists_are_bit_exact(emu, mixed, protein):
    offs, xyz, Q = _batch(mixed, protein)
    R = _run(emu, offs, xyz, Q, 48)
    assert np.array_equal(R["atom_sys"], np.repeat(np.arange(len(Q)), np.diff(offs)))
    pos = 0
    for s in range(len(Q)):
        a0, a1 = offs[s], offs[s + 1]
        D = O.distance_matrix(xyz[a0:a1])
        e_ref, _ = O.get_init_edges(xyz[a0:a1])
        near_ref = O.is_near_from_e(e_ref)
        assert np.all(R["q0"][a0:a1] == float(O.initial_charge(Q[s], a1 - a0)))
        inset = (D < 3.0) & ~np.eye(a1 - a0, dtype=bool)
        for i in range(a1 - a0):                                        # CSR rows: the e != 0 set, columns ascending, global indices
            r0, r1 = R["rowptr"][a0 + i], R["rowptr"][a0 + i + 1]
            assert np.array_equal(R["col"][r0:r1], a0 + np.nonzero(inset[i])[0]), (s, i)
        iu, ju = np.nonzero(np.triu(inset))
        k = len(iu)
        assert np.array_equal(R["pair_i"][pos:pos + k], a0 + iu) and np.array_equal(R["pair_j"][pos:pos + k], a0 + ju)
        assert np.array_equal(R["pair_D"][pos:pos + k], D[iu, ju])     # float64 distances, bit for bit
        assert np.array_equal(R["near"][pos:pos + k].astype(bool), near_ref[iu, ju])
        ee, er = R["e"][pos:pos + k].astype(np.float64), e_ref[iu, ju].astype(np.float64)
        assert np.all(np.abs(ee - er) <= np.spacing(np.abs(e_ref[iu, ju])).astype(np.float64))
        assert np.array_equal(R["e"][pos:pos + k].max(axis=1), e_ref[iu, ju].max(axis=1))   # the centre entries follow the reference formula exactly
        pos += k
    assert pos == R["P"]
    # reverse-edge index: every CSR entry points at its unordered pair
    rows = np.repeat(np.arange(offs[-1]), np.diff(R["rowptr"]))
    lo, hi = np.minimum(rows, R["col"]), np.maximum(rows, R["col"])
    assert np.array_equal(R["pair_i"][R["pid"]], lo) and np.array_equal(R["pair_j"][R["pid"]], hi)


def test_emulated_descriptor_coefficients(emu, mixed, protein):
    offs, xyz, Q = _batch(mixed, protein)
    R48 = _run(emu, offs, xyz, Q, 48)
    R16 = _run(emu, offs, xyz, Q, 16)
    assert np.array_equal(R16["near"], R48["near"]) and np.array_equal(R16["pair_D"], R48["pair_D"])
    c_ref = R48["e"].astype(np.float64) @ R16["B"]                       # c = B^T e in float64, rounded once
    assert np.abs(R16["e"] - c_ref).max() < 2e-7
    assert np.abs(R16["e"].astype(np.float64) @ R16["B"].T - R48["e"]).max() < 3e-7   # and B c gives e back (rank-16 family)
