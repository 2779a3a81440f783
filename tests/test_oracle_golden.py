"""Pins the CPU oracle to the reference's own shipped predictions."""
import numpy as np
import pytest

from oracle import epnn_oracle as O


def test_val_systems_match_shipped_predictions(weights, mixed, val871):
    """decay_model_weights, pad N=41: every 4th of the 871 validation systems (all 871 in the gpu suite)."""
    w = weights["decay_model_weights"]
    worst = 0.0
    for k in range(0, 871, 4):
        xyz, Z, Q = mixed.system(mixed.index[val871["names"][k]])
        q = O.forward_factorised(w, xyz, O.species_from_Z(Z, 9), Q, 41)
        n = len(Z)
        worst = max(worst, np.abs(q - val871["pred"][k, :n]).max())
        assert np.all(val871["pred"][k, n:] == 0)
    assert worst < 1.5e-6, worst          # survey measured 8.6e-7 over all 871


def test_literal_equals_factorised(weights, mixed):
    """The dense reference formulation and the factorised rewrite are the same function (all 3 checkpoints)."""
    for name, w in weights.items():
        for i in (0, 1500, 3000, 4300):
            xyz, Z, Q = mixed.system(i)
            sp = O.species_from_Z(Z, w.n_x)
            for npad in (len(Z), 41):
                a = O.forward_literal(w, xyz, sp, Q, npad)
                b = O.forward_factorised(w, xyz, sp, Q, npad)
                assert np.abs(a - b).max() < 1e-10, (name, i, npad)


def test_protein_matches_shipped_prediction(weights, protein):
    w = weights["decay_model_weights"]
    q = O.forward_factorised(w, protein["xyz"], O.species_from_Z(protein["Z"], 9), protein["Q"], None)
    assert np.abs(q - protein["preds"]).max() < 5e-6         # survey: 2.3e-6
    assert abs(q.sum() - 2.0) < 1e-6


def test_keras_input_wiring(weights, mixed):
    """make_model's un-tiling of the (N,N,.) inputs (charge_gn.py:382-384) reproduces forward_literal."""
    w = weights["model2_weights"]
    xyz, Z, Q = mixed.system(10)
    sp = O.species_from_Z(Z, 9)
    n, N = len(Z), len(Z) + 3
    x, h, q, e, mask = O._padded_inputs(w, xyz, sp, Q, N, np.float64)
    tile = lambda a: np.where(mask[:, :, None] > 0, np.broadcast_to(a[None, :, :], (N, N, a.shape[1])), 0.0)
    out = O.model_forward_keras_inputs(w, tile(h), e, tile(x), tile(q), mask)
    ref = O.forward_literal(w, xyz, sp, Q, N)
    assert np.abs(out[:n, 0] - ref).max() < 1e-12 and np.all(out[n:, 0] == 0)


def test_descriptor_properties(mixed):
    xyz, Z, Q = mixed.system(5)
    e, C = O.get_init_edges(xyz)
    assert e.dtype == np.float32 and np.array_equal(e, e.transpose(1, 0, 2))
    assert np.all(e[np.arange(len(Z)), np.arange(len(Z))] == 0)
    D = O.distance_matrix(xyz)
    assert np.all(e[D >= 3.0] == 0)
    near = O.is_near_from_e(e)
    assert np.array_equal(near, (D < 2.99396) & (D > 0) | ((D == 0) & ~np.eye(len(Z), dtype=bool)))
    mu = np.linspace(0.1, 3.0, 48)
    assert mu[-1] == 3.0


def test_pad_size_changes_result(weights, mixed):
    """SURVEY trap 1: padded atoms contribute to the unmasked message sum."""
    w = weights["model_weights"]
    xyz, Z, Q = mixed.system(100)
    sp = O.species_from_Z(Z, 10)
    a = O.forward_factorised(w, xyz, sp, Q, len(Z))
    b = O.forward_factorised(w, xyz, sp, Q, 41)
    assert np.abs(a - b).max() > 1e-3
