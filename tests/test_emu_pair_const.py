"""CPU warp emulation of the (GPU-unvalidated) pair-per-thread bundle kernels, epnn_b200/csrc/epnn_bundle_const.cu.

The kernel source is compiled for the host with -DEPNN_CPU_EMU (tools/emu/cuda_emu.h: every lane is a host thread,
shuffles / votes / __syncwarp through a per-warp barrier, shared memory a host buffer) and run on index lists built here
from the reference's own molecules exactly the way the CUDA prep kernels build them (pair lists sorted by i, per-bundle
far lists, species-compressed far lists, representatives).  Its outputs -- the message sums S of a message-passing step
and the per-pair transfers delta of an electron-passing pass -- are compared with a direct float64 evaluation of the
reference formulas (charge_gn.py:62-70, :101-116).  This checks the kernel's indexing, control flow, shared-memory
layout and summation logic before it ever reaches a GPU; it says nothing about performance."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from epnn_b200 import _capi
from oracle import epnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_bundle_const.so")
BUNDLE_ATOMS = 48


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    src = os.path.join(ROOT, "tools", "emu", "emu_bundle_const.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU", "-Wno-unknown-pragmas",
                           "-o", LIB, src])
    return C.CDLL(LIB)


def _build_lists(mixed, idx, n_x, npad, rng, equal_v_systems):
    offs, xyz, sp, _ = mixed.batch(idx, n_x)
    return _build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems)


def _build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems):
    """Everything the bundle kernels read, built on the host the way the CUDA prep kernels define it."""
    n = int(offs[-1])
    n_sys = len(offs) - 1
    atom_sys = np.repeat(np.arange(n_sys), np.diff(offs)).astype(np.int32)
    B = np.zeros((48, 16))
    assert _capi.load().epnn_rbf_basis(B.ctypes.data_as(C.c_void_p)) == 0
    pair_i, pair_j, near, coef = [], [], [], []
    rows = [[] for _ in range(n)]
    for s in range(n_sys):
        a0, a1 = offs[s], offs[s + 1]
        x = xyz[a0:a1]
        D = O.distance_matrix(x)
        e, _ = O.get_init_edges(x)
        for i in range(a1 - a0):
            for j in range(a1 - a0):
                if i != j and D[i, j] < 3.0:
                    rows[a0 + i].append(a0 + j)                       # CSR of the e != 0 set, columns ascending
                    if j > i:
                        pair_i.append(a0 + i); pair_j.append(a0 + j)
                        near.append(1 if e[i, j].max() > np.float32(1e-5) else 0)
                        coef.append((B.T @ e[i, j].astype(np.float64)).astype(np.float32))   # edge_desc_kernel: B^T e, rounded once
    P = len(pair_i)
    degU = np.bincount(np.array(pair_i, dtype=np.int64), minlength=n)
    ustart = np.concatenate([[0], np.cumsum(degU)]).astype(np.int32)
    # bundles: greedy runs of whole systems with <= BUNDLE_ATOMS atoms (run_chunk in epnn_api.cu)
    bundles, cur0, cur_n = [], -1, 0
    for s in range(n_sys):
        ns = offs[s + 1] - offs[s]
        if cur_n and cur_n + ns > BUNDLE_ATOMS:
            bundles.append((cur0, cur_n)); cur_n = 0
        if not cur_n:
            cur0 = int(offs[s])
        cur_n += int(ns)
    bundles.append((cur0, cur_n))
    atom_b0 = np.zeros(n, np.int32)
    for b0, bn in bundles:
        atom_b0[b0:b0 + bn] = b0
    # far lists (far_fill_kernel) and species-compressed far lists (far0_kernel)
    far_off, far_list, far0_off, far0_list, far0_w = [0], [], [0], [], []
    rep = np.zeros(n, np.int32)
    for i in range(n):
        s = atom_sys[i]
        a0, a1 = offs[s], offs[s + 1]
        b0 = atom_b0[i]
        rowset = set(rows[i])
        for j in range(a0, a1):
            if j not in rowset:
                far_list.append(((i - b0) << 8) | (j - b0))
        pad = npad[s] > a1 - a0
        if pad:
            far_list.append(((i - b0) << 8) | 0xFF)
        far_off.append(len(far_list))
        cnt = np.bincount(sp[a0:a1], minlength=16).astype(int)
        first = {}
        for j in range(a0, a1):
            first.setdefault(int(sp[j]), j)
        for j in rows[i]:
            cnt[sp[j]] -= 1
        rep[i] = first[int(sp[i])]
        for k in range(16):
            if cnt[k] > 0:
                far0_list.append(((i - b0) << 8) | (first[k] - b0)); far0_w.append(cnt[k])
        if pad:
            far0_list.append(((i - b0) << 8) | 0xFF); far0_w.append(0)
        far0_off.append(len(far0_list))
    # first-layer projections: random, with species-wise equal v rows in some systems (exercises the de-duplicated path)
    u = rng.normal(size=(n, 32)).astype(np.float32)
    v = rng.normal(size=(n, 32)).astype(np.float32)
    for s in equal_v_systems:
        table = rng.normal(size=(16, 32)).astype(np.float32)
        v[offs[s]:offs[s + 1]] = table[sp[offs[s]:offs[s + 1]]]
    return dict(offs=offs.astype(np.int32), sp=sp, n=n, n_sys=n_sys, atom_sys=atom_sys, rows=rows, P=P,
                pair_i=np.array(pair_i, np.int32), pair_j=np.array(pair_j, np.int32), near=np.array(near, np.uint8),
                coef=np.ascontiguousarray(np.array(coef, np.float32).reshape(P, 16)) if P else np.zeros((1, 16), np.float32), ustart=ustart,
                bundles=np.array(bundles, np.int32), far_off=np.array(far_off, np.int32), far_list=np.array(far_list, np.uint16),
                far0_off=np.array(far0_off, np.int32), far0_list=np.array(far0_list, np.uint16), far0_w=np.array(far0_w, np.uint8),
                rep=rep, u=u, v=v, npad=np.asarray(npad, np.int32))


def _run(emu, L, W, epn, dedup, n_warps):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    S = np.full((L["n"], 32), np.nan, np.float32)
    delta = np.full(L["P"], np.nan, np.float32)
    counter = np.zeros(1, np.int32)
    weights = np.concatenate([W["Cw"].ravel(), W["W2"].ravel(), W["b2"], W["x32"]]).astype(np.float32)
    emu.emu_bundle_const.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 13 + [C.c_int] + [C.c_void_p] * 7
    rc = emu.emu_bundle_const(int(epn), n_warps, p(weights), len(L["bundles"]), p(L["bundles"]), p(counter),
                              p(L["ustart"]), p(L["pair_i"]), p(L["pair_j"]), p(L["near"]), p(L["coef"]),
                              p(L["far_off"]), p(L["far_list"]), p(L["far0_off"]), p(L["far0_list"]), p(L["far0_w"]), p(L["rep"]), int(dedup),
                              p(L["atom_sys"]), p(L["offs"]), p(L["npad"]), p(L["u"]), p(L["v"]), p(S), p(delta))
    assert rc == 0
    return S, delta


def _weights(rng):
    return dict(Cw=(0.5 * rng.normal(size=(16, 32))).astype(np.float32), W2=(0.3 * rng.normal(size=(32, 32))).astype(np.float32),
                b2=(0.2 * rng.normal(size=32)).astype(np.float32), x32=rng.normal(size=32).astype(np.float32))


def _relu(x):
    return np.maximum(x, 0.0)


@pytest.mark.parametrize("dedup", [0, 1])
def test_emulated_gnn_bundle_kernel_matches_the_definition(emu, mixed, dedup):
    rng = np.random.default_rng(7)
    idx = mixed.usable(9)[[5, 40, 300, 1500, 1501, 2500, 3000, 3600, 4000]].tolist()
    sizes = np.array([mixed.offsets[i + 1] - mixed.offsets[i] for i in idx])
    npad = np.where(np.arange(len(idx)) % 3 == 1, sizes, 41)            # every third system without padding
    L = _build_lists(mixed, idx, 9, npad, rng, equal_v_systems=[0, 1, 2, 3, 6])
    assert len(L["bundles"]) >= 3
    W = _weights(rng)
    S, _ = _run(emu, L, W, epn=False, dedup=dedup, n_warps=3)
    Cw, W2, b2, b1 = (W[k].astype(np.float64) for k in ("Cw", "W2", "b2", "x32"))
    u, v = L["u"].astype(np.float64), L["v"].astype(np.float64)
    ce = {}
    for p in range(L["P"]):
        c = L["coef"][p].astype(np.float64) @ Cw
        ce[(L["pair_i"][p], L["pair_j"][p])] = c
        ce[(L["pair_j"][p], L["pair_i"][p])] = c
    ref = np.zeros((L["n"], 32))
    for i in range(L["n"]):
        s = L["atom_sys"][i]
        for j in range(L["offs"][s], L["offs"][s + 1]):                 # ALL columns of the system, the self pair included
            ref[i] += _relu(_relu(ce.get((i, j), 0.0) + u[i] + v[j]) @ W2 + b2)
        n_s = L["offs"][s + 1] - L["offs"][s]
        ref[i] += (npad[s] - n_s) * _relu(_relu(u[i] + b1) @ W2 + b2)   # padded atoms: a_j = 0, e = 0 -> v = b1
    assert np.isfinite(S).all()
    assert np.abs(S - ref).max() < 2e-5 * np.abs(ref).max(), np.abs(S - ref).max()


def test_emulated_epn_bundle_kernel_matches_the_definition(emu, mixed):
    rng = np.random.default_rng(8)
    idx = mixed.usable(9)[[7, 41, 301, 1600, 2501, 3001, 4100]].tolist()
    L = _build_lists(mixed, idx, 9, np.full(len(idx), 41), rng, equal_v_systems=[])
    L["near"][::7] = 0          # pairs with e != 0 that fail is_near are rare (2.994 A <= D < 3 A): plant some, the kernel only reads the flag
    W = _weights(rng)
    _, delta = _run(emu, L, W, epn=True, dedup=1, n_warps=2)
    Cw, W2, b2, w3 = (W[k].astype(np.float64) for k in ("Cw", "W2", "b2", "x32"))
    u, v = L["u"].astype(np.float64), L["v"].astype(np.float64)
    ref = np.zeros(L["P"])
    for p in range(L["P"]):
        i, j = L["pair_i"][p], L["pair_j"][p]
        c = L["coef"][p].astype(np.float64) @ Cw
        f_ij = _relu(_relu(c + u[i] + v[j]) @ W2 + b2) @ w3
        f_ji = _relu(_relu(c + u[j] + v[i]) @ W2 + b2) @ w3
        ref[p] = 0.5 * (f_ij - f_ji) * L["near"][p]                     # charge_gn.py:116
    assert L["near"].min() == 0 and L["near"].max() == 1               # both kinds of pair are present
    assert np.isfinite(delta).all()
    assert np.abs(delta - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), np.abs(delta - ref).max()


def _gnn_reference(L, W, npad):
    Cw, W2, b2, b1 = (W[k].astype(np.float64) for k in ("Cw", "W2", "b2", "x32"))
    u, v = L["u"].astype(np.float64), L["v"].astype(np.float64)
    ce = {}
    for p in range(L["P"]):
        c = L["coef"][p].astype(np.float64) @ Cw
        ce[(L["pair_i"][p], L["pair_j"][p])] = c
        ce[(L["pair_j"][p], L["pair_i"][p])] = c
    ref = np.zeros((L["n"], 32))
    for i in range(L["n"]):
        s = L["atom_sys"][i]
        for j in range(L["offs"][s], L["offs"][s + 1]):
            ref[i] += _relu(_relu(ce.get((i, j), 0.0) + u[i] + v[j]) @ W2 + b2)
        ref[i] += (npad[s] - (L["offs"][s + 1] - L["offs"][s])) * _relu(_relu(u[i] + b1) @ W2 + b2)
    return ref


@pytest.mark.parametrize("dedup", [0, 1])
def test_emulated_gnn_edge_cases(emu, mixed, dedup):
    """A single atom, two atoms beyond the cutoff (no e != 0 pair at all), a bundle filled to exactly 48 atoms (last
    staged row next to the pad row), a bundle without any padded system, more warps than bundles."""
    rng = np.random.default_rng(11)
    ok = mixed.usable(9)
    sizes = np.diff(mixed.offsets)[ok]
    i29 = int(ok[np.nonzero(sizes == 29)[0][0]])
    i19 = int(ok[np.nonzero(sizes == 19)[0][0]])
    o2, x2, s2, _ = mixed.batch([i29, i19], 9)
    o3, x3, s3, _ = mixed.batch([int(ok[3])], 9)
    offs = np.concatenate([o2, [49, 51], 51 + o3[1:]]).astype(np.int32)                 # 29 | 19 | 1 | 2 | n
    xyz = np.concatenate([x2, [[0.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [5.0, 0.0, 0.0]], x3]).astype(np.float32)
    sp = np.concatenate([s2, [1], [0, 3], s3]).astype(np.int32)
    npad = np.array([29, 19, 1, 7, 41], np.int32)                      # systems 0, 1, 2 unpadded: the first bundle has no pad slot
    L = _build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems=[0, 1, 2, 3])
    assert int(L["bundles"][0][1]) == 48 and len(L["bundles"]) == 2
    W = _weights(rng)
    S, _ = _run(emu, L, W, epn=False, dedup=dedup, n_warps=8)
    ref = _gnn_reference(L, W, npad)
    assert np.isfinite(S).all()
    assert np.abs(S - ref).max() < 2e-5 * np.abs(ref).max(), np.abs(S - ref).max()
