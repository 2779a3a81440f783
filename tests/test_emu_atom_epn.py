"""CPU warp emulation of the per-atom kernel (epnn_atom.cu: folded update MLP, first-layer projections, fixed-order
charge reduction over the CSR rows) and of the large-system electron-passing pair kernel (epnn_epn.cu) -- the unmodified
CUDA source against float64 evaluations of what each mode is documented to compute."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from emu_common import build_lists_raw, csr, epn_reference, large_system_tables, relu, weights
from oracle import epnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_atom_epn.so")
UPDATE, QUPDATE, PROJECT, OUTPUT, FIRST, WRITE_H = 1, 2, 4, 8, 16, 32      # epnn_internal.cuh


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU", "-Wno-unknown-pragmas",
                           "-o", LIB, os.path.join(ROOT, "tools", "emu", "emu_atom_epn.cpp")])
    lib = C.CDLL(LIB)
    lib.emu_atom_kernel.argtypes = [C.c_int] * 4 + [C.c_void_p] * 18
    lib.emu_atom_const_kernel.argtypes = [C.c_int] * 4 + [C.c_void_p] * 18
    lib.emu_atom_mma_kernel.argtypes = [C.c_int] * 4 + [C.c_void_p] * 18
    lib.emu_epn_pair_kernel.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _lists(protein, mixed, rng):
    """A 70-atom protein cut (large: several S planes) between two small molecules; 87+ atoms = three 32-atom tiles."""
    xs, zs, offs = [], [], [0]
    for part in (int(mixed.usable(9)[3]), None, int(mixed.usable(9)[20])):
        if part is None:
            x, z = protein["xyz"][50:120], protein["Z"][50:120]
        else:
            x, z, _ = mixed.system(part)
        xs.append(x); zs.append(O.species_from_Z(z, 9)); offs.append(offs[-1] + len(z))
    offs = np.array(offs, np.int32)
    npad = np.array([41, 75, 29], np.int32)
    L = build_lists_raw(offs, np.concatenate(xs).astype(np.float32), np.concatenate(zs).astype(np.int32), npad, rng, [])
    return L, npad


# FP32 SIMT warp-tile kernel; experimental atom-per-thread variant; warp-level tensor kernel (3xTF32, the FP32 default)
@pytest.mark.parametrize("kernel", ["emu_atom_kernel", "emu_atom_const_kernel", "emu_atom_mma_kernel"])
def test_emulated_atom_kernel_modes(emu, protein, mixed, kernel):
    rng = np.random.default_rng(21)
    L, npad = _lists(protein, mixed, rng)
    n = L["n"]
    rowptr, col = csr(L)
    _, _, _, pid, _ = large_system_tables(L)
    f32 = lambda *shape: rng.normal(size=shape).astype(np.float32)
    HG, cb, g, U2, c2, U3, c3 = 0.3 * f32(64, 32), f32(32), 0.05 * f32(32), 0.3 * f32(32, 32), f32(32), 0.3 * f32(32, 48), f32(48)
    Pf, Aq, Ax = 0.3 * f32(32, 64), f32(64), f32(16, 64)
    wu = np.concatenate([x.ravel() for x in (HG, cb, g, U2, c2, U3, c3)])
    wp = np.concatenate([x.ravel() for x in (Pf, Aq, Ax)])
    nsplit = 3
    S = f32(nsplit, n, 32)
    l2_prev = np.abs(f32(n, 32))
    q0 = rng.normal(size=n)
    delta = f32(L["P"])
    sys_of = L["atom_sys"]
    large = (np.diff(L["offs"])[sys_of] > 48)
    Ssum = np.where(large[:, None], S.astype(np.float64).sum(axis=0), S[0].astype(np.float64))     # small systems: plane 0 only
    npf = npad[sys_of].astype(np.float64)

    def run(mode, h_is_zero=0, q=q0):
        l2 = l2_prev.copy(); h = np.full((n, 48), np.nan, np.float32); u = np.full((n, 32), np.nan, np.float32); v = u.copy()
        qd = q.copy(); qo = np.full(n, np.nan, np.float32); qo64 = np.full(n, np.nan)
        assert getattr(emu, kernel)(mode, h_is_zero, n, nsplit, _p(wu), _p(wp), _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]), _p(L["sp"]),
                                   _p(S), _p(h), _p(l2), _p(rowptr), _p(col), _p(pid), _p(delta), _p(qd), _p(u), _p(v), _p(qo), _p(qo64)) == 0
        return dict(l2=l2, h=h, u=u, v=v, q=qd, qo=qo, qo64=qo64)

    def uv_of(l2, q):
        out = (l2 @ Pf.astype(np.float64) if l2 is not None else 0.0) + Ax.astype(np.float64)[L["sp"]] + q[:, None] * Aq.astype(np.float64)
        return out[:, :32], out[:, 32:]

    close = lambda a, b, tol=2e-5: np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())
    # (A) first projection: h = 0, no product -- u | v = Ax[species] + q Aq
    r = run(PROJECT, h_is_zero=1)
    ur, vr = uv_of(None, q0)
    assert close(r["u"], ur) and close(r["v"], vr)
    # (B) first update (l2_prev ignored: h = 0) and (C) later update, the last one also materialising h = U3^T l2 + c3
    for mode, prev in ((UPDATE | PROJECT | FIRST, np.zeros((n, 32))), (UPDATE | PROJECT | WRITE_H, l2_prev.astype(np.float64))):
        r = run(mode)
        l1 = relu(np.concatenate([prev, Ssum], axis=1) @ HG.astype(np.float64) + cb + npf[:, None] * g.astype(np.float64))
        l2 = relu(l1 @ U2.astype(np.float64) + c2)
        ur, vr = uv_of(l2, q0)
        assert close(r["l2"], l2) and close(r["u"], ur, 5e-5) and close(r["v"], vr, 5e-5)
        if mode & WRITE_H:
            assert close(r["h"], l2 @ U3.astype(np.float64) + c3, 5e-5)
    # (D) charge update + projection for the next pass, (E) charge update + output: q_i += sum over the row of +/- delta
    qn = q0.copy()
    for i in range(n):
        for k in range(rowptr[i], rowptr[i + 1]):
            d = float(delta[pid[k]])
            qn[i] += d if col[k] > i else -d
    r = run(QUPDATE | PROJECT)
    ur, vr = uv_of(l2_prev.astype(np.float64), qn)
    assert np.abs(r["q"] - qn).max() < 1e-12 and close(r["u"], ur, 5e-5) and close(r["v"], vr, 5e-5)
    r = run(QUPDATE | OUTPUT)
    assert np.abs(r["qo64"] - qn).max() < 1e-12 and np.array_equal(r["qo"], qn.astype(np.float32))
    assert abs(r["qo64"].sum() - q0.sum()) < 1e-9                     # every transfer enters twice with opposite signs


def test_emulated_large_epn_pair_kernel(emu, protein, mixed):
    rng = np.random.default_rng(22)
    L, _ = _lists(protein, mixed, rng)
    L["near"][::5] = 0
    W = weights(rng)
    wts = np.concatenate([W["Cw"].ravel(), W["W2"].ravel(), W["b2"], W["x32"]]).astype(np.float32)
    delta = np.full(L["P"], np.nan, np.float32)
    assert emu.emu_epn_pair_kernel(_p(wts), L["P"], _p(L["pair_i"]), _p(L["pair_j"]), _p(L["near"]), _p(L["coef"]),
                                   _p(L["atom_sys"]), _p(L["offs"]), _p(L["u"]), _p(L["v"]), _p(delta)) == 0
    ref = epn_reference(L, W)
    in_large = np.diff(L["offs"])[L["atom_sys"][L["pair_i"]]] > 48
    assert in_large.any() and (~in_large).any()
    assert np.isnan(delta[~in_large]).all()                            # pairs of small systems belong to the bundle kernel
    assert np.abs(delta[in_large] - ref[in_large]).max() < 2e-5 * max(1.0, np.abs(ref).max())
