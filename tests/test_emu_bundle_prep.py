"""CPU warp emulation of the fused list building for chunks of small systems (epnn_bundle_prep.cu: one warp per bundle, a
48-bit neighbour mask per row): every list it writes -- CSR, pair ids, local rows, unordered pairs and their
descriptors, far list, species-compressed far list, representatives, offsets -- against the host construction that follows the
definitions of the general kernels (tests/emu_common.py) and the oracle's distances."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from emu_common import build_lists, csr, large_system_tables
from epnn_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_bundle_prep.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU", "-Wno-unknown-pragmas",
                           "-o", LIB, os.path.join(ROOT, "tools", "emu", "emu_bundle_prep.cpp")])
    lib = C.CDLL(LIB)
    lib.emu_bundle_prep.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 28 + [C.c_int]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed,n_sys,pad", [(1, 40, 41), (2, 25, 0), (3, 60, 29)])
def test_fused_lists_match_the_definitions(emu, mixed, seed, n_sys, pad):
    rng = np.random.default_rng(seed)
    idx = sorted(rng.choice(mixed.usable(9), n_sys, replace=False).tolist())
    sizes = np.array([len(mixed.system(i)[1]) for i in idx])
    npad = np.maximum(sizes, pad).astype(np.int32) if pad else sizes.astype(np.int32)       # pad 0: no pad slot anywhere
    if pad == 29:
        npad = np.where(np.arange(n_sys) % 3 == 0, sizes, np.maximum(sizes, 41)).astype(np.int32)   # mixed: some systems unpadded
    L = build_lists(mixed, idx, 9, npad, rng, [])
    offs, xyz, sp, _ = mixed.batch(idx, 9)
    n, nb = L["n"], len(L["bundles"])
    cap = 64 * n
    i32 = lambda k: np.full(k, -7, np.int32)
    tot, deg, degU, rowptr, ustart, far_off, far0_off, rep, atom_b0, bnat = i32(4), i32(n), i32(n), i32(n + 1), i32(n + 1), i32(n + 1), i32(n + 1), i32(n), i32(n), i32(n)
    col, pid, pair_i, pair_j = i32(cap), i32(cap), i32(cap), i32(cap)
    rowl = np.full(cap, 255, np.uint8); coef = np.full((cap, 16), np.nan, np.float32); near = np.full(cap, 255, np.uint8)
    mu, B = np.zeros(48), np.zeros((48, 16))
    lib = _capi.load()
    assert lib.epnn_rbf_centers(_p(mu)) == 0 and lib.epnn_rbf_basis(_p(B)) == 0
    far_list = np.zeros(cap, np.uint16); far0_list = np.zeros(cap, np.uint16); far0_w = np.full(cap, 255, np.uint8)
    bundles = np.ascontiguousarray(L["bundles"], np.int32)
    assert emu.emu_bundle_prep(nb, n, _p(bundles), _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]), _p(L["sp"]), _p(np.ascontiguousarray(xyz, np.float32)),
                               _p(tot), _p(deg), _p(degU), _p(rowptr), _p(ustart), _p(far_off), _p(far0_off), _p(rep), _p(atom_b0), _p(bnat),
                               _p(col), _p(pid), _p(rowl), _p(pair_i), _p(pair_j), _p(mu), _p(B), _p(coef), _p(near), _p(far_list), _p(far0_list), _p(far0_w), cap) == 0
    ref_rowptr, ref_col = csr(L)
    _, _, _, ref_pid, ref_deg = large_system_tables(L)
    nnz, P = int(ref_rowptr[-1]), L["P"]
    assert tot.tolist() == [nnz, P, len(L["far_list"]), len(L["far0_list"])]
    assert np.array_equal(rowptr, ref_rowptr) and np.array_equal(col[:nnz], ref_col[:nnz]) and np.array_equal(deg, ref_deg)
    assert np.array_equal(ustart, L["ustart"]) and np.array_equal(pair_i[:P], L["pair_i"]) and np.array_equal(pair_j[:P], L["pair_j"])
    assert np.array_equal(degU, np.diff(L["ustart"]))
    assert np.array_equal(pid[:nnz], ref_pid[:nnz])
    b0_of = np.zeros(n, np.int32)
    for b0, bn in L["bundles"]:
        b0_of[b0:b0 + bn] = b0
        assert bnat[b0] == bn
    assert np.array_equal(atom_b0, b0_of)
    rows_of = np.repeat(np.arange(n), np.diff(ref_rowptr))
    assert np.array_equal(rowl[:nnz], (rows_of - b0_of[rows_of]).astype(np.uint8))
    assert np.array_equal(far_off, L["far_off"]) and np.array_equal(far_list[:len(L["far_list"])], L["far_list"])
    assert np.array_equal(far0_off, L["far0_off"]) and np.array_equal(far0_list[:len(L["far0_list"])], L["far0_list"])
    assert np.array_equal(far0_w[:len(L["far0_w"])], L["far0_w"]) and np.array_equal(rep, L["rep"])
    # the pair's distance is evaluated by the descriptor kernel itself (one thread per pair): near flags bit-exact, coefficients
    # = B^T e of the float32 descriptors to float32 round-off (build_lists follows the oracle's get_init_edges)
    assert np.array_equal(near[:P], L["near"]) and np.abs(coef[:P] - L["coef"]).max() < 2e-7
