"""Synthetic workloads of BASELINE.json configs 4 and 5, generated from the committed
QM9 / protein fixtures (there is no network for datasets).  Deterministic in (seed, n)."""
from __future__ import annotations

import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_Z9 = np.array([1, 6, 7, 8, 9, 16, 17, 35])
_Z10 = np.array([1, 6, 7, 8, 9, 15, 16, 17, 35])


def species_from_Z(Z, n_x):
    table = _Z9 if n_x == 9 else _Z10
    lut = np.full(64, -1, np.int32)
    lut[table] = np.arange(len(table), dtype=np.int32)
    sp = lut[np.asarray(Z, dtype=np.int64)]
    if (sp < 0).any():
        raise KeyError("element not in the table selected by n_x")
    return sp


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
                     np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
                     np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1)], 1)


def qm9_pool():
    d = np.load(os.path.join(_GOLDEN, "mixed.npz"))
    names = [str(x) for x in d["names"]]
    idx = [i for i, n in enumerate(names) if n.startswith("dsgdb9nsd")]
    offs = d["offsets"]
    return idx, offs, d["xyz"], d["Z"]


def qm9_shaped(n_mol: int, n_x: int = 9, seed: int = 0, first: int = 0):
    """Config 4: ``n_mol`` molecules drawn with replacement from the 1338 QM9 molecules of data/mixed
    (<= 29 atoms), each randomly rotated and jittered by N(0, 0.01 A), Q = 0.  Molecule k of the stream
    depends only on (seed, first + k), so ranks can generate disjoint shards independently.
    Returns (offsets int32, xyz float32 (n,3), species int32, Q float32)."""
    idx, offs, xyz_all, Z_all = qm9_pool()
    idx = np.asarray(idx)
    sizes_all = (offs[1:] - offs[:-1])[idx]
    block = 65536
    out_off, out_xyz, out_sp = [np.zeros(1, np.int64)], [], []
    k = first
    end = first + n_mol
    while k < end:
        b0 = (k // block) * block
        rng = np.random.default_rng([seed, b0 // block])
        pick = rng.integers(0, len(idx), size=block)
        rot = _random_rotations(rng, block)
        lo, hi = k - b0, min(end, b0 + block) - b0
        pick_s, rot_s = pick[lo:hi], rot[lo:hi]
        sizes = sizes_all[pick_s]
        o = np.concatenate([[0], np.cumsum(sizes)])
        mol_of_atom = np.repeat(np.arange(hi - lo), sizes)
        within = np.arange(o[-1]) - o[mol_of_atom]
        src = offs[idx[pick_s]][mol_of_atom] + within
        jit_rng = np.random.default_rng([seed, b0 // block, 1])
        jitter_all = jit_rng.normal(scale=0.01, size=(int(sizes_all[pick].sum()), 3))
        j0 = int(sizes_all[pick[:lo]].sum())
        x = xyz_all[src].astype(np.float64)
        cen = np.add.reduceat(x, o[:-1], axis=0) / sizes[:, None]
        x = np.einsum("aij,aj->ai", rot_s[mol_of_atom], x - cen[mol_of_atom]) + jitter_all[j0:j0 + o[-1]]
        out_xyz.append(x.astype(np.float32))
        out_sp.append(species_from_Z(Z_all[src], n_x))
        out_off.append(out_off[-1][-1] + o[1:])
        k = b0 + hi
    offsets = np.concatenate(out_off)
    if offsets[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 atoms in one shard")
    return (offsets.astype(np.int32), np.concatenate(out_xyz), np.concatenate(out_sp).astype(np.int32),
            np.zeros(n_mol, np.float32))


def protein():
    d = np.load(os.path.join(_GOLDEN, "protein.npz"))
    return d["xyz"].astype(np.float32), d["Z"], np.float32(d["Q"])


def protein_like(n_atoms: int, n_x: int = 9, seed: int = 1):
    """Config 5: Galectin-3C tiled on a 3-D grid (spacing = bounding box + 1.5 A so copies interact across
    faces), one random rotation per copy, truncated to ``n_atoms``; Q = +2 per (started) copy."""
    xyz, Z, Q = protein()
    n1 = len(Z)
    copies = -(-n_atoms // n1)
    side = int(np.ceil(copies ** (1 / 3)))
    rng = np.random.default_rng(seed)
    cen = xyz.mean(0)
    x0 = xyz.astype(np.float64) - cen
    r = np.linalg.norm(x0, axis=1).max()
    spacing = 2 * r / np.sqrt(3) + 1.5          # rotated copies overlap mildly at the faces; reject clashes below
    from scipy.spatial import cKDTree
    out = []
    placed = []          # (centre, lazily built KD-tree holder, coordinates)
    c = 0
    for ix in range(side):
        for iy in range(side):
            for iz in range(side):
                if c >= copies:
                    break
                for _ in range(64):
                    R = _random_rotations(rng, 1)[0]
                    x = x0 @ R.T + np.array([ix, iy, iz]) * spacing
                    ok = True
                    cx = x.mean(0)
                    for prev in placed[-side * side - side - 1:]:
                        if np.linalg.norm(prev[0] - cx) < 2 * r:
                            if prev[1][0] is None:
                                prev[1][0] = cKDTree(prev[2])
                            if prev[1][0].query(x, k=1, distance_upper_bound=0.96)[0].min() < 0.96:
                                ok = False
                                break
                    if ok:
                        break
                placed.append((x.mean(0), [None], x))
                out.append(x)
                c += 1
    x = np.concatenate(out)[:n_atoms].astype(np.float32)
    sp = np.tile(species_from_Z(Z, n_x), copies)[:n_atoms]
    return np.array([0, n_atoms], np.int32), x, sp.astype(np.int32), np.array([2.0 * copies], np.float32)
