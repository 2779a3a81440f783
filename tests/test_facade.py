"""The reference-named facade (epnn_b200.charge_gn / epnn_b200.infer): CPU-side behaviour + GPU parity of the
dense Keras-shaped call against the oracle's literal restatement of make_model (charge_gn.py:369-391)."""
import os

import numpy as np
import pytest

from oracle import epnn_oracle as O

XYZ_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xyz") + "/"


def test_tables_match_reference_conventions():
    from epnn_b200 import charge_gn
    assert charge_gn.elem_dict == {'H': 0, 'C': 1, 'N': 2, 'O': 3, 'F': 4, 'P': 5, 'S': 6, 'Cl': 7, 'Br': 8}
    assert charge_gn.atom_num_dict['Br'] == 35 and charge_gn.atom_num_dict['P'] == 15


def test_make_model_argument_checks(golden_dir):
    from epnn_b200 import charge_gn
    m = charge_gn.make_model([32, 32], 48, 5, 9, 41)
    with pytest.raises(RuntimeError):
        m([None] * 5)                                   # no weights loaded: there is no random-init path
    with pytest.raises(NotImplementedError):
        charge_gn.make_model([64, 64], 48, 5, 9, 41)
    m3 = charge_gn.make_model([32, 32], 48, 5, 9, 41)
    with pytest.raises(ValueError):                      # model2_weights has T=3: Keras would fail to restore too
        m3.load_weights(os.path.join(golden_dir, "checkpoints", "model2_weights"))


def test_get_init_edges_rejects_what_the_reference_rejects():
    from epnn_b200 import charge_gn
    with pytest.raises(SystemExit):                      # charge_gn.py:140-145
        charge_gn.get_init_edges(np.zeros((2, 3), np.float32), np.array([1, 2]), num=48)
    with pytest.raises(NotImplementedError):
        charge_gn.get_init_edges(np.zeros((2, 3), np.float32), np.array([]), num=32)


def _dense_inputs(w, xyz, sp, Q, N, rng=None):
    """(h, e, x, q, mask) for one system, tiled like gen_padded_init_state (charge_gn.py:342-364)."""
    x, h, q, e, mask = O._padded_inputs(w, xyz, sp, Q, N, np.float64)
    if rng is not None:                                  # non-trivial hidden state: exercises the h rows of the first layer
        h[:len(sp)] = rng.normal(scale=0.3, size=(len(sp), 48))
    tile = lambda a: np.where(mask[:, :, None] > 0, np.broadcast_to(a[None, :, :], (N, N, a.shape[1])), 0.0)
    return tile(h), e, tile(x), tile(q), mask


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["decay_model_weights", "model2_weights", "model_weights"])
def test_dense_call_matches_literal_oracle(weights, mixed, golden_dir, name):
    from epnn_b200 import charge_gn
    w = weights[name]
    rng = np.random.default_rng(3)
    idx = [7, 1500, 4200]
    N = 41
    batch = []
    for i in idx:
        xyz, Z, Q = mixed.system(i)
        batch.append(_dense_inputs(w, xyz, O.species_from_Z(Z, w.n_x), Q, N, rng))
    h, e, x, q, mask = (np.stack([b[k] for b in batch]) for k in range(5))
    model = charge_gn.make_model([32, 32], 48, w.T, w.n_x, N, precision=64)
    model.load_weights(os.path.join(golden_dir, "checkpoints", name))
    out = model([h, e, x, q, mask])
    assert out.shape == (3, N, 1) and out.dtype == np.float32
    for b in range(3):
        ref = O.model_forward_keras_inputs(w, h[b], e[b], x[b], q[b], mask[b])
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(out[b] - ref).max() < 2e-6 * scale, (name, b)
    # FP32 kernels, default checkpoint: still inside the north-star tolerance
    if name == "decay_model_weights":
        m32 = charge_gn.make_model([32, 32], 48, w.T, w.n_x, N)
        m32.load_weights(os.path.join(golden_dir, "checkpoints", name))
        out32 = m32([h, e, x, q, mask[..., None]])           # (B,N,N,1) mask like the Keras Input (charge_gn.py:380)
        assert np.abs(out32 - out).max() < 1e-5


@pytest.mark.gpu
def test_dense_call_arbitrary_mask_and_asymmetric_e(weights, golden_dir):
    """Nothing about the inputs is assumed: ragged mask, asymmetric e, generic x."""
    from epnn_b200 import charge_gn
    w = weights["model2_weights"]
    rng = np.random.default_rng(8)
    N = 9
    h = rng.normal(scale=0.2, size=(1, N, N, 48))
    e = np.abs(rng.normal(scale=0.3, size=(1, N, N, 48))).astype(np.float32) * (rng.random((1, N, N, 1)) > 0.4)
    x = rng.normal(size=(1, N, N, 9))
    q = rng.normal(scale=0.1, size=(1, N, N, 1))
    mask = (rng.random((1, N, N)) > 0.3).astype(np.float64)
    mask[0, :, 4] = 0                                    # an atom nobody sees: divide_no_nan -> 0
    model = charge_gn.make_model([32, 32], 48, 3, 9, N, precision=64)
    model.load_weights(os.path.join(golden_dir, "checkpoints", "model2_weights"))
    out = model([h, e, x, q, mask])
    ref = O.model_forward_keras_inputs(w, h[0], e[0], x[0], q[0], mask[0])
    assert np.abs(out[0] - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
def test_gen_padded_init_state_and_cli(weights, golden_dir, tmp_path, monkeypatch, capsys):
    from epnn_b200 import charge_gn, infer, xyzio
    x, h, q, e, Q, y, mask, names = charge_gn.gen_padded_init_state(XYZ_DIR, 48, 48, n_elems=9)
    systems = xyzio.read_directory(XYZ_DIR)
    S, N = len(systems), max(s.n for s in systems)
    assert x.shape == (S, N, N, 9) and h.shape == (S, N, N, 48) and q.shape == (S, N, N, 1)
    assert e.shape == (S, N, N, 48) and y.shape == (S, N, 1) and mask.shape == (S, N, N) and x.dtype == np.float64
    assert list(names) == [s.name for s in systems]
    w = weights["decay_model_weights"]
    for i, s in enumerate(systems):
        n = s.n
        ref_e, _ = O.get_init_edges(s.xyz)
        assert np.abs(e[i, :n, :n] - ref_e).max() <= 1e-7 and np.all(e[i, n:] == 0) and np.all(e[i, :, n:] == 0)
        assert np.all(mask[i, :n, :n] == 1) and mask[i].sum() == n * n
        assert np.allclose(q[i, :n, :n, 0], np.float32(np.float32(s.Q) / n))
        assert np.array_equal(x[i, 0, :n], O.features(O.species_from_symbols(s.symbols, 9), 9))
        if s.labels is not None:
            assert np.allclose(y[i, :n, 0], s.labels)
    # CLI, fast packed path and literal dense path, against the oracle at the same pad size
    monkeypatch.chdir(tmp_path)
    ck = os.path.join(golden_dir, "checkpoints", "decay_model_weights")
    fast = infer.main(["--path", XYZ_DIR, "--weights", ck, "--out", "fast.npy", "--repeats", "2"])
    dense = infer.main(["--path", XYZ_DIR, "--weights", ck, "--dense"])
    out = capsys.readouterr().out
    assert "avg inference time" in out and "avg feature time" in out and "MAE vs labels" in out
    assert os.path.exists("test_names.npy") and np.load("fast.npy").shape == (S, 2, N, 1)
    for i, s in enumerate(systems):
        ref = O.forward_factorised(w, s.xyz, O.species_from_symbols(s.symbols, 9), s.Q, N)
        assert np.abs(fast[i, 0, :s.n, 0] - ref).max() < 1e-5
        assert np.abs(dense[i, 0, :s.n, 0] - ref).max() < 1e-5
        assert np.all(fast[i, :, s.n:] == 0)
