"""The library's native xyz ingest (epnn_xyz_load / epnn_xyz_parse_text, host code only) against the Python parser
that restates the reference's loop (charge_gn.py:309-330).  Runs without a GPU."""
import os

import numpy as np
import pytest

from epnn_b200 import elements, xyzio

XYZ_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xyz")


def _python_pack(paths, n_x):
    return xyzio.pack([xyzio.read_xyz(p) for p in paths], n_x)


@pytest.mark.parametrize("n_x", [9, 10])
def test_native_equals_python_on_reference_files(n_x):
    paths = sorted(os.path.join(XYZ_DIR, f) for f in os.listdir(XYZ_DIR) if f.endswith(".xyz"))
    a = xyzio.load_packed(paths, n_x, threads=3)
    b = _python_pack(paths, n_x)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y.reshape(x.shape))
    names, offs, xyz, sp, Q = xyzio.read_directory_packed(XYZ_DIR, n_x, sort=True)
    assert names == [os.path.basename(p)[:-4] for p in paths] and np.array_equal(offs, a[0])


def test_dialect_details_and_rounding():
    text = ("3\n-1 2 ignored tokens\n"
            "Cl   1.0000000149011612e+00 -2.5   3e-1  extra columns 7 8\n"
            "\n"                                     # blank lines are skipped
            "H\t0.1\t0.2\t0.30000001192092896\r\n"     # tabs, CRLF
            "Br 1e10 -0.0 12345678.9")               # no trailing newline
    offs, xyz, sp, Q = xyzio.parse_packed(text, 9)
    ref = xyzio.parse_xyz_text(text)
    assert np.array_equal(offs, [0, 3]) and Q[0] == np.float32(-1)
    assert np.array_equal(xyz, ref.xyz) and xyz.dtype == np.float32          # float64 parse, one rounding to float32
    assert np.array_equal(sp, elements.species_index(ref.symbols, 9))
    assert np.array_equal(xyzio.parse_packed(text, 10)[2], elements.species_index(ref.symbols, 10))


def test_errors_match_the_reference_behaviour(tmp_path):
    with pytest.raises(KeyError):                     # charge_gn.py:326-327: dict lookup of an unknown element
        xyzio.parse_packed("1\n0 1\nXx 0 0 0\n", 9)
    with pytest.raises(KeyError):                     # P only exists in the 10-wide table
        xyzio.parse_packed("1\n0 1\nP 0 0 0\n", 9)
    assert xyzio.parse_packed("1\n0 1\nP 0 0 0\n", 10)[2][0] == 5
    with pytest.raises(ValueError):
        xyzio.parse_packed("1\n0 1\nH 0 0\n", 9)      # missing coordinate
    with pytest.raises(ValueError):
        xyzio.parse_packed("1\nfoo\nH 0 0 0\n", 9)    # charge is not a number
    with pytest.raises(ValueError):
        xyzio.parse_packed("1\n0 1\n", 9)             # no atoms
    for bad in ("nan", "inf", "-Infinity", "0x1p3", "0X10", "1,5"):      # non-finite, hex floats, locale decimal commas
        with pytest.raises(ValueError):
            xyzio.parse_packed(f"1\n0 1\nH 0 {bad} 0\n", 9)
    for bad in ("nan", "inf"):
        with pytest.raises(ValueError):
            xyzio.parse_xyz_text(f"1\n0 1\nH 0 {bad} 0\n")
        with pytest.raises(ValueError):
            xyzio.parse_packed(f"1\n{bad}\nH 0 0 0\n", 9)
    with pytest.raises(FileNotFoundError):
        xyzio.load_packed([str(tmp_path / "missing.xyz")], 9)


def test_many_files_in_parallel_keep_their_order(tmp_path):
    rng = np.random.default_rng(0)
    syms = ["H", "C", "N", "O", "F", "S", "Cl", "Br"]
    paths = []
    for k in range(300):
        n = int(rng.integers(1, 30))
        lines = [str(n), f"{int(rng.integers(-2, 3))} 1"]
        for _ in range(n):
            c = rng.normal(scale=3.0, size=3)
            lines.append(f"{syms[int(rng.integers(0, 8))]} {c[0]:.8f} {c[1]:.8f} {c[2]:.8f}")
        p = tmp_path / f"m{k:04d}.xyz"
        p.write_text("\n".join(lines) + "\n")
        paths.append(str(p))
    a = xyzio.load_packed(paths, 9, threads=8)
    b = _python_pack(paths, 9)
    for x, y in zip(a, b):
        assert np.array_equal(x, y.reshape(x.shape))
    c = xyzio.load_packed(paths[::-1], 9, threads=1)
    assert np.array_equal(np.diff(c[0]), np.diff(a[0])[::-1])
