import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CKPTS = ("decay_model_weights", "model_weights", "model2_weights")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_usable():
    """True if a CUDA device can actually be used (the driver is asked through torch; no kernel runs)."""
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:      # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them (the product has no CPU
    fallback, so they cannot pass there).  On a GPU box nothing is skipped: a missing library then fails loudly."""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device: epnn_b200 has no CPU fallback (run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def weights():
    from epnn_b200.checkpoint import load_weights
    return {n: load_weights(os.path.join(GOLDEN, "checkpoints", n)) for n in CKPTS}


class Mixed:
    """The packed data/mixed set (tests/golden/mixed.npz)."""

    def __init__(self):
        d = np.load(os.path.join(GOLDEN, "mixed.npz"))
        self.names = [str(x) for x in d["names"]]
        self.offsets = d["offsets"]
        self.xyz = d["xyz"]
        self.Z = d["Z"]
        self.Q = d["Q"]
        self.index = {n: i for i, n in enumerate(self.names)}

    def usable(self, n_x):
        """Indices of systems whose elements all exist in the table chosen by n_x (the 9-wide one has no P)."""
        if n_x == 10:
            return np.arange(len(self.names))
        hasP = np.add.reduceat((self.Z == 15).astype(np.int64), self.offsets[:-1]) > 0
        return np.nonzero(~hasP)[0]

    def system(self, i):
        a0, a1 = self.offsets[i], self.offsets[i + 1]
        return self.xyz[a0:a1], self.Z[a0:a1], self.Q[i]

    def batch(self, idx, n_x):
        from oracle.epnn_oracle import species_from_Z
        offs = [0]
        xyz, sp, Q = [], [], []
        for i in idx:
            x, z, q = self.system(i)
            xyz.append(x)
            sp.append(species_from_Z(z, n_x))
            Q.append(q)
            offs.append(offs[-1] + len(z))
        return (np.array(offs, np.int32), np.concatenate(xyz).astype(np.float32), np.concatenate(sp).astype(np.int32),
                np.array(Q, np.float32))


@pytest.fixture(scope="session")
def mixed():
    return Mixed()


@pytest.fixture(scope="session")
def val871():
    d = np.load(os.path.join(GOLDEN, "val871.npz"))
    return {"names": [str(x) for x in d["names"]], "pred": d["pred"], "lab": d["lab"]}


@pytest.fixture(scope="session")
def protein():
    d = np.load(os.path.join(GOLDEN, "protein.npz"))
    return {"xyz": d["xyz"], "Z": d["Z"], "Q": np.float32(d["Q"]), "preds": d["preds"].reshape(-1)}


@pytest.fixture(scope="session")
def engines(weights):
    """One CUDA engine per checkpoint (gpu tests only)."""
    from epnn_b200.engine import Engine
    made = {}

    def get(name, precision=32):
        key = (name, precision)
        if key not in made:
            made[key] = Engine(weights[name], device=0, precision=precision)
        return made[key]

    yield get
    for e in made.values():
        e.close()
