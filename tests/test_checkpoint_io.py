"""Host I/O: TF tensor-bundle reader, xyz dialect, element tables (CPU)."""
import os

import numpy as np
import pytest

from epnn_b200 import checkpoint, elements, xyzio


def test_checkpoint_shapes(weights):
    # (T, n_x, K, parameter count)
    expect = {"decay_model_weights": (5, 9, 164, 74037), "model_weights": (5, 10, 166, 74677),
              "model2_weights": (3, 9, 164, 46515)}
    for name, (T, n_x, K, n) in expect.items():
        w = weights[name]
        assert (w.T, w.n_x, w.K) == (T, n_x, K)
        assert w.packed().size == n
        assert len(w.msg) == T and len(w.pas) == T
        for m in w.pas:      # final pass bias is exactly zero in all shipped checkpoints
            assert np.all(m.b[-1] == 0)


def test_step_aliasing(golden_dir):
    """The last step's MLP is stored under message_fn / pass_fn, not message_fns/<T-1> (charge_gn.py:61,99)."""
    t = checkpoint.read_bundle(os.path.join(golden_dir, "checkpoints", "decay_model_weights"))
    keys = [k for k in t if "message_fns/" in k]
    assert {int(k.split("message_fns/")[1].split("/")[0]) for k in keys} == {0, 1, 2, 3}
    assert any(k.startswith("layer_with_weights-0/message_fn/") for k in t)
    w = checkpoint.load_weights(os.path.join(golden_dir, "checkpoints", "decay_model_weights"))
    last = t["layer_with_weights-0/message_fn/layer_set/0/kernel/.ATTRIBUTES/VARIABLE_VALUE"]
    assert np.array_equal(w.msg[4].W[0], last)


def test_crc_detects_corruption(golden_dir, tmp_path):
    src = os.path.join(golden_dir, "checkpoints")
    for f in os.listdir(src):
        if f.startswith("model2_weights"):
            data = bytearray(open(os.path.join(src, f), "rb").read())
            if f.endswith("data-00001-of-00002"):
                data[1000] ^= 0xFF
            open(tmp_path / f, "wb").write(bytes(data))
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_weights(str(tmp_path / "model2_weights"))
    checkpoint.load_weights(str(tmp_path / "model2_weights"), verify_crc=False)   # still parseable without CRC


def test_bad_magic(tmp_path):
    open(tmp_path / "x.index", "wb").write(b"\0" * 100)
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.read_index(str(tmp_path / "x.index"))


def test_crc32c_known_answer():
    assert checkpoint.crc32c(b"123456789") == 0xE3069283


def test_xyz_dialect(golden_dir):
    d = os.path.join(golden_dir, "xyz")
    systems = xyzio.read_directory(d, sort=True)
    assert len(systems) == 7
    for s in systems:
        assert s.xyz.dtype == np.float32 and s.xyz.shape == (s.n, 3)
        assert s.labels is not None and s.labels.shape == (s.n,)
    assert {float(s.Q) for s in systems} >= {0.0, 1.0}
    s = xyzio.parse_xyz_text("3\n-1 2 ignored\nO 0 0 0 extra\nH 0 0 1\nH 0 1 0\n", "w")
    assert s.Q == np.float32(-1) and s.symbols == ["O", "H", "H"]
    off, xyz, sp, Q = xyzio.pack([s], 9)
    assert off.tolist() == [0, 3] and sp.tolist() == [3, 0, 0] and Q.tolist() == [-1.0]


def test_element_tables():
    assert elements.symbols_for(9) == ('H', 'C', 'N', 'O', 'F', 'S', 'Cl', 'Br')       # infer.py:22-30
    assert elements.symbols_for(10)[5] == 'P'                                             # charge_gn.py:19-28
    x = elements.features(np.array([0, 5]), 9)
    assert x[0, 0] == 1 and x[0, 1] == 1 and x[1, 0] == 16 and x[1, 6] == 1 and x.sum() == 19
    with pytest.raises(KeyError):
        elements.species_index(["P"], 9)          # P is not in the 9-wide table (reference raises KeyError too)
    with pytest.raises(KeyError):
        elements.species_index(["Xx"], 10)


@pytest.mark.parametrize("name", ["decay_model_weights", "model_weights", "model2_weights"])
def test_save_weights_reproduces_the_shipped_files_byte_for_byte(golden_dir, tmp_path, name):
    """checkpoint.save_weights (replaces model.save_weights, charge_gn.py:462): re-saving an unmodified shipped checkpoint
    must give the reference's own .index and .data shards back, byte for byte (SSTable layout, restart points, block
    CRCs, BundleEntryProto bytes, tensor bytes and CRCs)."""
    src = os.path.join(golden_dir, "checkpoints", name)
    w = checkpoint.load_weights(src)
    dst = str(tmp_path / "resaved")
    checkpoint.save_weights(w, dst, src)
    files = sorted(f for f in os.listdir(os.path.dirname(src)) if f.startswith(name + "."))
    assert len(files) >= 2
    for f in files:
        a = open(os.path.join(os.path.dirname(src), f), "rb").read()
        b = open(dst + f[len(name):], "rb").read()
        assert a == b, f


def test_save_weights_round_trip_of_modified_weights(golden_dir, tmp_path):
    src = os.path.join(golden_dir, "checkpoints", "model2_weights")
    w = checkpoint.load_weights(src)
    rng = np.random.default_rng(3)
    for mlp in w.msg + w.pas + [w.upd]:
        for k in range(len(mlp.W)):
            mlp.W[k] = (mlp.W[k] + rng.normal(size=mlp.W[k].shape)).astype(np.float32)
            mlp.b[k] = (mlp.b[k] - 1.0).astype(np.float32)
    dst = str(tmp_path / "edited")
    checkpoint.save_weights(w, dst, src)
    back = checkpoint.load_weights(dst)                     # verifies every block and tensor CRC on the way
    assert np.array_equal(back.packed(), w.packed())
    assert not np.array_equal(back.packed(), checkpoint.load_weights(src).packed())
    # the Keras object graph (what TensorFlow restores by) is carried over untouched
    assert checkpoint.read_index(dst + ".index")["_CHECKPOINTABLE_OBJECT_GRAPH"] == \
        checkpoint.read_index(src + ".index")["_CHECKPOINTABLE_OBJECT_GRAPH"]


def test_save_weights_rejects_a_template_of_another_architecture(golden_dir, tmp_path):
    w = checkpoint.load_weights(os.path.join(golden_dir, "checkpoints", "model2_weights"))      # T = 3
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.save_weights(w, str(tmp_path / "x"), os.path.join(golden_dir, "checkpoints", "decay_model_weights"))   # T = 5
    w10 = checkpoint.load_weights(os.path.join(golden_dir, "checkpoints", "model_weights"))     # n_x = 10
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.save_weights(w10, str(tmp_path / "y"), os.path.join(golden_dir, "checkpoints", "decay_model_weights"))
