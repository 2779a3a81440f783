"""CPU warp emulation of the DEFAULT bundle kernels (epnn_b200/csrc/epnn_bundle.cu) and of the kernels that build their
lists: the unmodified CUDA source, compiled for the host with -DEPNN_CPU_EMU (tools/emu/cuda_emu.h), against
 (a) lists built independently in numpy (far lists, species-compressed far lists, representatives), and
 (b) a float64 evaluation of the reference formulas (charge_gn.py:62-70, :101-116) for the message sums S of a
     message-passing step -- with and without the far-column de-duplication -- and the transfers delta of an
     electron-passing pass.
The same kernels are validated on a B200 by tests/test_gpu_parity.py; this CPU replica exists so that changes to the
warp-synchronous code (scatter_sorted, tile permutation, far lists, prefetch pipeline) can be checked before any
GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from emu_common import build_lists, csr, epn_reference, gnn_reference, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libemu_bundle.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DEPNN_CPU_EMU", "-Wno-unknown-pragmas",
                           "-o", LIB, os.path.join(ROOT, "tools", "emu", "emu_bundle.cpp")])
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _device_lists(emu, L):
    """far / far0 / perm lists from the real prep kernels, run thread by thread."""
    n, P = L["n"], L["P"]
    rowptr, col = csr(L)
    out = dict(atom_b0=np.zeros(n + 1, np.int32), bundle_nat=np.zeros(n + 1, np.int32), far_off=np.zeros(n + 1, np.int32),
               far_list=np.zeros(50 * n + 2, np.uint16), far0_off=np.zeros(n + 1, np.int32), far0_list=np.zeros(17 * n + 2, np.uint16),
               far0_w=np.zeros(17 * n + 16, np.uint8), rep=np.zeros(n + 1, np.int32), perm_j=np.zeros(P + 16, np.uint8))
    n_far, n_far0 = C.c_int(0), C.c_int(0)
    emu.emu_bundle_lists.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 18 + [C.POINTER(C.c_int)] * 2
    rc = emu.emu_bundle_lists(n, len(L["bundles"]), _p(L["bundles"]), P, _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]), _p(L["sp"]),
                              _p(rowptr), _p(col), _p(L["ustart"]), _p(L["pair_i"]), _p(L["pair_j"]),
                              _p(out["atom_b0"]), _p(out["bundle_nat"]), _p(out["far_off"]), _p(out["far_list"]),
                              _p(out["far0_off"]), _p(out["far0_list"]), _p(out["far0_w"]), _p(out["rep"]), _p(out["perm_j"]),
                              C.byref(n_far), C.byref(n_far0))
    assert rc == 0
    out["n_far"], out["n_far0"] = n_far.value, n_far0.value
    return out


def _run_kernel(emu, L, D, W, epn, dedup):
    S = np.full((L["n"], 32), np.nan, np.float32)
    delta = np.full(max(L["P"], 1), np.nan, np.float32)
    counter = np.zeros(1, np.int32)
    wts = np.concatenate([W["Cw"].ravel(), W["W2"].ravel(), W["b2"], W["x32"]]).astype(np.float32)
    emu.emu_bundle_kernel.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 14 + [C.c_int] + [C.c_void_p] * 7
    rc = emu.emu_bundle_kernel(int(epn), _p(wts), len(L["bundles"]), _p(L["bundles"]), _p(counter),
                               _p(L["ustart"]), _p(L["pair_i"]), _p(L["pair_j"]), _p(L["near"]), _p(L["coef"]), _p(D["perm_j"]),
                               _p(D["far_off"]), _p(D["far_list"]), _p(D["far0_off"]), _p(D["far0_list"]), _p(D["far0_w"]), _p(D["rep"]),
                               int(dedup), _p(L["atom_sys"]), _p(L["offs"]), _p(L["npad"]), _p(L["u"]), _p(L["v"]), _p(S), _p(delta))
    assert rc == 0
    return S, delta[:L["P"]]


def _case(mixed, rng, equal_v):
    idx = mixed.usable(9)[[5, 40, 300, 1500, 1501, 2500, 3000, 3600, 4000]].tolist()
    sizes = np.array([mixed.offsets[i + 1] - mixed.offsets[i] for i in idx])
    npad = np.where(np.arange(len(idx)) % 3 == 1, sizes, 41)
    return build_lists(mixed, idx, 9, npad, rng, equal_v_systems=equal_v), npad


def test_prep_kernels_build_the_lists_numpy_builds(emu, mixed):
    L, _ = _case(mixed, np.random.default_rng(1), [])
    D = _device_lists(emu, L)
    assert np.array_equal(D["far_off"][:L["n"] + 1], L["far_off"]) and D["n_far"] == len(L["far_list"])
    assert np.array_equal(D["far_list"][:D["n_far"]], L["far_list"])
    assert np.array_equal(D["far0_off"][:L["n"] + 1], L["far0_off"]) and D["n_far0"] == len(L["far0_list"])
    assert np.array_equal(D["far0_list"][:D["n_far0"]], L["far0_list"])
    assert np.array_equal(D["far0_w"][:D["n_far0"]], L["far0_w"])
    assert np.array_equal(D["rep"][:L["n"]], L["rep"])
    # tile permutation: inside every 32-pair tile of a bundle, perm_j ranks the pairs by (j, slot)
    for b0, bn in L["bundles"]:
        p0, p1 = L["ustart"][b0], L["ustart"][b0 + bn]
        for tb in range(p0, p1, 32):
            js = L["pair_j"][tb:min(tb + 32, p1)]
            assert np.array_equal(np.argsort(np.argsort(js, kind="stable"), kind="stable"), D["perm_j"][tb:tb + len(js)])


@pytest.mark.parametrize("dedup", [0, 1])
def test_emulated_default_gnn_bundle_kernel(emu, mixed, dedup):
    rng = np.random.default_rng(7)
    L, npad = _case(mixed, rng, [0, 1, 2, 3, 6])
    D = _device_lists(emu, L)
    W = weights(rng)
    S, _ = _run_kernel(emu, L, D, W, epn=False, dedup=dedup)
    ref = gnn_reference(L, W, npad)
    assert np.isfinite(S).all()
    assert np.abs(S - ref).max() < 2e-5 * np.abs(ref).max(), np.abs(S - ref).max()


def test_emulated_default_epn_bundle_kernel(emu, mixed):
    rng = np.random.default_rng(8)
    L, _ = _case(mixed, rng, [])
    L["near"][::7] = 0
    D = _device_lists(emu, L)
    W = weights(rng)
    _, delta = _run_kernel(emu, L, D, W, epn=True, dedup=1)
    ref = epn_reference(L, W)
    assert np.isfinite(delta).all()
    assert np.abs(delta - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), np.abs(delta - ref).max()
