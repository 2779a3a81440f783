"""Shared helpers of the CPU warp-emulation tests (tests/test_emu_*.py): index lists built on the host the way the CUDA
prep kernels define them, random weights, float64 evaluation of the reference formulas."""
import ctypes as C

import numpy as np

from epnn_b200 import _capi
from oracle import epnn_oracle as O

BUNDLE_ATOMS = 48


def build_lists(mixed, idx, n_x, npad, rng, equal_v_systems):
    offs, xyz, sp, _ = mixed.batch(idx, n_x)
    return build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems)


def build_lists_raw(offs, xyz, sp, npad, rng, equal_v_systems):
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
        if ns > BUNDLE_ATOMS:                                          # large systems take the row-group kernels
            if cur_n:
                bundles.append((cur0, cur_n)); cur_n = 0
            continue
        if cur_n and cur_n + ns > BUNDLE_ATOMS:
            bundles.append((cur0, cur_n)); cur_n = 0
        if not cur_n:
            cur0 = int(offs[s])
        cur_n += int(ns)
    if cur_n:
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
        if a1 - a0 > BUNDLE_ATOMS:                                     # far_count_kernel / far0_kernel: nothing for large systems
            far_off.append(len(far_list)); far0_off.append(len(far0_list)); rep[i] = i
            continue
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
                bundles=np.array(bundles, np.int32).reshape(-1, 2), far_off=np.array(far_off, np.int32), far_list=np.array(far_list, np.uint16),
                far0_off=np.array(far0_off, np.int32), far0_list=np.array(far0_list, np.uint16), far0_w=np.array(far0_w, np.uint8),
                rep=rep, u=u, v=v, npad=np.asarray(npad, np.int32))


def weights(rng):
    return dict(Cw=(0.5 * rng.normal(size=(16, 32))).astype(np.float32), W2=(0.3 * rng.normal(size=(32, 32))).astype(np.float32),
                b2=(0.2 * rng.normal(size=32)).astype(np.float32), x32=rng.normal(size=32).astype(np.float32))


def relu(x):
    return np.maximum(x, 0.0)


def gnn_reference(L, W, npad):
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
            ref[i] += relu(relu(ce.get((i, j), 0.0) + u[i] + v[j]) @ W2 + b2)
        ref[i] += (npad[s] - (L["offs"][s + 1] - L["offs"][s])) * relu(relu(u[i] + b1) @ W2 + b2)
    return ref




def epn_reference(L, W):
    Cw, W2, b2, w3 = (W[k].astype(np.float64) for k in ("Cw", "W2", "b2", "x32"))
    u, v = L["u"].astype(np.float64), L["v"].astype(np.float64)
    ref = np.zeros(L["P"])
    for p in range(L["P"]):
        i, j = L["pair_i"][p], L["pair_j"][p]
        c = L["coef"][p].astype(np.float64) @ Cw
        f_ij = relu(relu(c + u[i] + v[j]) @ W2 + b2) @ w3
        f_ji = relu(relu(c + u[j] + v[i]) @ W2 + b2) @ w3
        ref[p] = 0.5 * (f_ij - f_ji) * L["near"][p]                     # charge_gn.py:116
    return ref


def csr(L):
    """rowptr / col of the e != 0 set (both directions, columns ascending) from the per-row lists."""
    deg = np.array([len(r) for r in L["rows"]], np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = np.array([j for r in L["rows"] for j in r] + [0], np.int32)
    return rowptr, col


def large_system_tables(L):
    """Row groups of the large systems (rg_fill_kernel), rgl_off (exclusive scan of the per-system counts), pair id and
    degree per CSR entry / atom -- what gnn_pair_kernel<LARGE> reads besides the CSR."""
    offs = L["offs"]
    n_sys = L["n_sys"]
    cnt = [((offs[s + 1] - offs[s] + 3) // 4) if offs[s + 1] - offs[s] > BUNDLE_ATOMS else 0 for s in range(n_sys)]
    rgl_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    rg = [i for s in range(n_sys) if cnt[s] for i in range(offs[s], offs[s + 1], 4)]
    pid_of = {(int(i), int(j)): p for p, (i, j) in enumerate(zip(L["pair_i"], L["pair_j"]))}
    pid = np.array([pid_of[(min(i, j), max(i, j))] for i, r in enumerate(L["rows"]) for j in r] + [0], np.int32)
    deg = np.array([len(r) for r in L["rows"]], np.int32)
    return np.array(rg + [0], np.int32), len(rg), rgl_off, pid, deg
