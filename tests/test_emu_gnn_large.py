"""CPU warp emulation of the large-system message kernel (gnn_pair_kernel<float, LARGE>, epnn_b200/csrc/epnn_gnn.cu) with
and without the exact de-duplication of far columns, including the species tables (sp_tab_fill: __match_any_sync +
integer atomics) and the per-step equality check -- the unmodified CUDA source against a float64 evaluation of the
reference's unmasked message sum (charge_gn.py:66-70)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from emu_common import build_lists_raw, csr, gnn_reference, large_system_tables, weights
from oracle import epnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_gnn.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU", "-Wno-unknown-pragmas",
                           "-o", LIB, os.path.join(ROOT, "tools", "emu", "emu_gnn.cpp")])
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _system(protein, mixed, n_x, rng, ninth_species):
    """150-, 61- and 49-atom cuts of the protein (large), with a small molecule in between (ignored by this kernel)."""
    cuts = [(0, 150), None, (400, 461), (900, 949)]
    offs, xyz, sp = [0], [], []
    for c in cuts:
        if c is None:
            x, z, _ = mixed.system(int(mixed.usable(9)[10]))
        else:
            x, z = protein["xyz"][c[0]:c[1]], protein["Z"][c[0]:c[1]]
        xyz.append(x); sp.append(O.species_from_Z(z, n_x)); offs.append(offs[-1] + len(z))
    sp = np.concatenate(sp).astype(np.int32)
    if ninth_species:                                                   # 10-wide table: P (5) and Br (8) present
        pick = rng.permutation(150)
        sp[pick[:12]] = 8; sp[pick[12:18]] = 5
    return np.array(offs, np.int32), np.concatenate(xyz).astype(np.float32), sp


@pytest.mark.parametrize("nsplit,n_x,ninth", [(1, 9, False), (3, 9, False), (2, 10, True)])
def test_emulated_large_gnn_kernel_with_and_without_dedup(emu, protein, mixed, nsplit, n_x, ninth):
    rng = np.random.default_rng(13)
    offs, xyz, sp = _system(protein, mixed, n_x, rng, ninth)
    sizes = np.diff(offs)
    npad = np.array([sizes[0] + 6, 41, sizes[2], sizes[3] + 1], np.int32)
    L = build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems=[0, 3])      # system 2 keeps random v rows: no collapse there
    rowptr, col = csr(L)
    rg, n_rg, rgl_off, pid, deg = large_system_tables(L)
    assert n_rg == 38 + 16 + 13
    W = weights(rng)
    wts = np.concatenate([W["Cw"].ravel(), W["W2"].ravel(), W["b2"], W["x32"]]).astype(np.float32)
    ref = gnn_reference(L, W, npad)
    large = np.concatenate([np.arange(offs[s], offs[s + 1]) for s in (0, 2, 3)])
    n_tab = n_rg // 8 + 2
    emu.emu_gnn_large_step.argtypes = ([C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 12 +
                                       [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] * 2)
    for stamp in (0, 2):
        S = np.zeros((nsplit, L["n"], 32), np.float32)
        tab = np.zeros(n_tab * 32, np.int32); st = np.zeros(n_tab * 2, np.int32); rows = np.zeros(1, np.uint64)
        rc = emu.emu_gnn_large_step(_p(wts), L["n"], L["n_sys"], n_rg, _p(rg), nsplit, 0,
                                    _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]), _p(L["sp"]), _p(rgl_off),
                                    _p(rowptr), _p(col), _p(pid), _p(deg), _p(L["coef"]), _p(L["u"]), _p(L["v"]),
                                    stamp, n_x - 1, n_tab, _p(tab), _p(st), _p(rows), _p(S), 0, 1)
        assert rc == 0
        tot = S.astype(np.float64).sum(axis=0)                          # the per-atom kernel adds the planes in this order
        err = np.abs(tot[large] - ref[large]).max() / np.abs(ref[large]).max()
        assert err < 2e-5, (stamp, nsplit, err)
        if stamp:
            assert int(rows[0]) == sizes[0] + sizes[3]                  # systems 0 and 3 collapse, system 2 does not
            t0 = tab[(rgl_off[0] >> 3) * 32:(rgl_off[0] >> 3) * 32 + 32]
            assert np.array_equal(t0[:16], np.bincount(sp[:150], minlength=16))        # atoms per species
            assert all(t0[16 + k] == np.nonzero(sp[:150] == k)[0][0] for k in range(16) if t0[k])   # first atom per species
            if nsplit > 1:                                              # collapsed systems leave the other planes at zero
                assert np.abs(S[1:, :150]).max() == 0.0


def test_emulated_sharded_large_gnn_is_bit_identical(emu, protein, mixed):
    """epnn_set_shard splits the row-group x plane units of this kernel into contiguous ranges, zero-fills the rest and
    all-reduces: the partial buffers of the ranks must add up to the single-rank buffer BIT FOR BIT (each element is
    written by exactly one rank), with the de-duplication on."""
    rng = np.random.default_rng(14)
    offs, xyz, sp = _system(protein, mixed, 9, rng, False)
    sizes = np.diff(offs)
    npad = np.array([sizes[0] + 6, 41, sizes[2], sizes[3] + 1], np.int32)
    L = build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems=[0])
    rowptr, col = csr(L)
    rg, n_rg, rgl_off, pid, deg = large_system_tables(L)
    W = weights(rng)
    wts = np.concatenate([W["Cw"].ravel(), W["W2"].ravel(), W["b2"], W["x32"]]).astype(np.float32)
    n_tab, nsplit = n_rg // 8 + 2, 3
    emu.emu_gnn_large_step.argtypes = ([C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 12 +
                                       [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] * 2)

    def run(rank, world):
        S = np.zeros((nsplit, L["n"], 32), np.float32)
        tab = np.zeros(n_tab * 32, np.int32); st = np.zeros(n_tab * 2, np.int32); rows = np.zeros(1, np.uint64)
        assert emu.emu_gnn_large_step(_p(wts), L["n"], L["n_sys"], n_rg, _p(rg), nsplit, 0, _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]),
                                      _p(L["sp"]), _p(rgl_off), _p(rowptr), _p(col), _p(pid), _p(deg), _p(L["coef"]), _p(L["u"]), _p(L["v"]),
                                      1, 8, n_tab, _p(tab), _p(st), _p(rows), _p(S), rank, world) == 0
        return S

    single = run(0, 1)
    for world in (2, 3):
        parts = [run(r, world) for r in range(world)]
        assert sum((p != 0).astype(int) for p in parts).max() <= 1          # disjoint support
        assert np.array_equal(sum(parts), single)
