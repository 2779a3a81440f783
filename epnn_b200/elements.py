"""Element tables / per-atom feature conventions of the reference.

Two conventions exist upstream and the checkpoint decides which applies (SURVEY.md trap 2):

* n_x == 10: ``charge_gn.py:9-28``  -> H C N O F P S Cl Br   (x = [Z, onehot9])
* n_x ==  9: ``infer.py:13-30``     -> H C N O F S Cl Br     (x = [Z, onehot8]); the convention
  ``decay_model_weights`` / ``model2_weights`` were trained with.

Unknown element symbols raise ``KeyError`` exactly like the reference's dict lookups
(``charge_gn.py:326-327``).
"""
from __future__ import annotations

import numpy as np

ATOMIC_NUMBER = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'P': 15, 'S': 16, 'Cl': 17, 'Br': 35}

_SYMBOLS = {
    10: ('H', 'C', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br'),
    9: ('H', 'C', 'N', 'O', 'F', 'S', 'Cl', 'Br'),
}


def symbols_for(n_x: int):
    if n_x not in _SYMBOLS:
        raise ValueError(f"no element table for n_x={n_x} (expected 9 or 10)")
    return _SYMBOLS[n_x]


def species_index(symbols, n_x: int) -> np.ndarray:
    """Map element symbols to int32 indices into the table chosen by ``n_x`` (KeyError if unknown)."""
    table = {s: i for i, s in enumerate(symbols_for(n_x))}
    return np.array([table[s] for s in symbols], dtype=np.int32)


def z_table(n_x: int) -> np.ndarray:
    return np.array([ATOMIC_NUMBER[s] for s in symbols_for(n_x)], dtype=np.float32)


def features(species: np.ndarray, n_x: int) -> np.ndarray:
    """x_i = [Z, onehot(species)]  (reference ``charge_gn.py:325-328``), float32 (n, n_x)."""
    species = np.asarray(species)
    x = np.zeros((species.shape[0], n_x), dtype=np.float32)
    x[:, 0] = z_table(n_x)[species]
    x[np.arange(species.shape[0]), species + 1] = 1.0
    return x
