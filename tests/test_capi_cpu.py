"""The C-ABI library loads without a GPU and exports every symbol include/epnn_b200.h declares (CPU)."""
import ctypes as C
import os
import re

import numpy as np

from epnn_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "epnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epnn_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = _declared()
    assert set(names) == set(_capi.SIGNATURES), (names, sorted(_capi.SIGNATURES))


def test_every_declared_symbol_is_exported():
    lib = _capi.load()
    for n in _declared():
        assert getattr(lib, n) is not None


def test_version_and_rbf_centers():
    lib = _capi.load()
    assert b"sm_100a" in lib.epnn_version()
    mu = np.zeros(48)
    assert lib.epnn_rbf_centers(mu.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(mu, np.linspace(0.1, 3.0, 48))          # bit-exact with charge_gn.py:123


def test_create_rejects_bad_arguments_before_touching_cuda():
    lib = _capi.load()
    h = C.c_void_p()
    w = np.zeros(10, np.float32)
    assert lib.epnn_create(0, 5, 9, w.ctypes.data_as(C.c_void_p), 10, C.byref(h)) == -1
    assert b"expected 74037" in lib.epnn_last_error(None)
    assert lib.epnn_create(0, 5, 11, w.ctypes.data_as(C.c_void_p), 10, C.byref(h)) == -1
    assert lib.epnn_create(0, 5, 9, None, 74037, C.byref(h)) == -1


def test_product_never_imports_oracle():
    """The product package must not route through the oracle or any CPU fallback."""
    pkg = os.path.join(ROOT, "epnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_descriptor_basis_is_orthonormal_and_complete_to_1e9():
    """The FP32 kernels carry B^T e (16 numbers) per pair instead of e (48): B must be orthonormal and the family of
    radial descriptors e(D) (charge_gn.py:148-161) must lie in span(B) to far below float32 round-off for EVERY D."""
    lib = _capi.load()
    B = np.zeros((48, 16))
    assert lib.epnn_rbf_basis(B.ctypes.data_as(C.c_void_p)) == 0
    assert np.abs(B.T @ B - np.eye(16)).max() < 1e-12
    mu = np.linspace(0.1, 3.0, 48)
    D = np.concatenate([np.linspace(0.0, 3.0, 100001)[:-1], np.random.default_rng(1).uniform(0, 3, 50000)])
    Cc = (np.cos(np.pi * D / 3.0) + 1.0) / 2.0
    E = Cc[:, None] * np.exp(-2.0 * (D[:, None] - mu[None]) ** 2)
    assert np.abs(E - (E @ B) @ B.T).max() < 1e-9


def test_stats_struct_matches_the_header():
    """The ctypes mirror of epnn_stats must list the header's fields in the same order with the same C types."""
    text = open(os.path.join(ROOT, "include", "epnn_b200.h")).read()
    body = re.search(r"typedef struct epnn_stats \{(.*?)\} epnn_stats;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int64_t|int32_t|float)\s+([a-z0-9_]+)\s*;", body)
    ctype = {"int64_t": C.c_int64, "int32_t": C.c_int32, "float": C.c_float}
    size = {"int64_t": 8, "int32_t": 4, "float": 4}
    assert [(n, ctype[t]) for t, n in fields] == list(_capi.Stats._fields_)
    packed = sum(size[t] for t, _ in fields)
    assert C.sizeof(_capi.Stats) == (packed + 7) // 8 * 8          # no internal padding; tail padded to the int64 alignment


def test_bench_real_data_workloads_are_the_reference_sets():
    """bench.py --workload qm9_test / ssi: the 1338 QM9 molecules / 2979 SSI dimers of data/mixed, packed consistently."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    for which, n_x, n_sys, n_atoms in (("qm9_test", 10, 1338, 24033), ("ssi", 9, 2979, 65904)):
        offs, xyz, sp, Q, n = bench.real_set(which, n_x)
        assert (n, int(offs[-1])) == (n_sys, n_atoms)
        assert xyz.shape == (n_atoms, 3) and sp.shape == (n_atoms,) and Q.shape == (n_sys,)
        assert np.diff(offs).min() >= 3 and np.diff(offs).max() <= 41 and sp.min() >= 0 and sp.max() < n_x - 1
    offs, _, _, Q, n = bench.real_set("ssi", 9, limit=16)
    assert n == 16 and len(offs) == 17
    assert set(np.unique(bench.real_set("ssi", 9)[3]).tolist()) == {-2.0, -1.0, 0.0, 1.0, 2.0}
