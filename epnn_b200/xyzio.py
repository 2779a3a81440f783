"""Reader for the reference's xyz dialect (``charge_gn.py:309-330``).

line 1: atom count (ignored: atoms = every line after line 2)
line 2: first token = net charge Q (float32); remaining tokens ignored
line 3+: ``Elem x y z [ignored columns]``; coordinates are rounded to float32 like the reference.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class System:
    name: str
    symbols: List[str]
    xyz: np.ndarray            # (n, 3) float32
    Q: np.float32
    labels: Optional[np.ndarray] = None   # (n,) float32 MBIS labels if <stem>.npy exists

    @property
    def n(self) -> int:
        return len(self.symbols)


def parse_xyz_text(text: str, name: str = "") -> System:
    lines = text.splitlines()
    if len(lines) < 3:
        raise ValueError(f"{name}: xyz file needs at least one atom line")
    Q = np.float32(lines[1].strip().split()[0])
    symbols, coords = [], []
    for line in lines[2:]:
        data = line.split()
        if not data:            # the reference would IndexError on a blank line; tolerate a trailing one
            continue
        symbols.append(data[0])
        coords.append([data[1], data[2], data[3]])
    xyz = np.array(coords, dtype=np.float32)
    return System(name=name, symbols=symbols, xyz=xyz, Q=Q)


def read_xyz(path: str) -> System:
    with open(path, "r") as f:
        s = parse_xyz_text(f.read(), os.path.basename(path)[:-4])
    lab = path[:-4] + ".npy"
    if os.path.exists(lab):
        s.labels = np.array(np.load(lab), dtype=np.float32).reshape(-1)
    return s


def read_directory(path: str, sort: bool = False) -> List[System]:
    """All ``*.xyz`` of a directory in ``os.listdir`` order (reference ``charge_gn.py:301``)."""
    names = os.listdir(path)
    if sort:
        names = sorted(names)
    return [read_xyz(os.path.join(path, n)) for n in names if n.endswith(".xyz")]


def pack(systems: Sequence[System], n_x: int):
    """Concatenate systems into the SoA arrays the C-ABI takes: offsets, xyz, species, Q."""
    from .elements import species_index
    offsets = np.zeros(len(systems) + 1, dtype=np.int32)
    for i, s in enumerate(systems):
        offsets[i + 1] = offsets[i] + s.n
    xyz = np.concatenate([s.xyz for s in systems], axis=0).astype(np.float32) if systems else np.zeros((0, 3), np.float32)
    species = np.concatenate([species_index(s.symbols, n_x) for s in systems]) if systems else np.zeros(0, np.int32)
    Q = np.array([s.Q for s in systems], dtype=np.float32)
    return offsets, np.ascontiguousarray(xyz), np.ascontiguousarray(species, dtype=np.int32), Q
