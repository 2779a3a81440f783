"""The FP32 kernel sets ("pair_const" option) against the oracle and against each other.

  2 (default since round 2): row-run GNN bundle kernel (epnn_bundle_run.cu) + pair-per-thread EPN bundle kernel
     (epnn_bundle_const.cu) + row-per-thread far kernel for large systems (epnn_gnn_far_const.cu) + warp-tile per-atom kernel
  1: pair-per-thread kernels everywhere (GNN bundle kernel with the 32 x 33 transpose, per-atom variant)
  0: the round-1 warp-tile kernels

All three are plain FP32 and evaluate the same formulas; only the mapping of the work onto the warp (and therefore the --
fixed -- order of the additions into S) differs, so the bar is the same for each: FP32 tolerances against the float64
oracle, 871 shipped predictions within 1e-5, hidden state against the oracle for the live checkpoints, charge
conservation, bitwise reproducibility, agreement with the default set to FP32 round-off.  First validated on a B200 in
round 2 (profiles/r02/call01_pair_const_tests_and_ab.log)."""
import numpy as np
import pytest

from oracle import epnn_oracle as O

pytestmark = [pytest.mark.gpu]

TOL_FP32 = {"decay_model_weights": 2e-6, "model2_weights": 1e-5, "model_weights": 2e-4}     # tests/test_gpu_parity.py: measured floors


def _engine(weights, name, kset=1):
    from epnn_b200.engine import Engine
    eng = Engine(weights[name], device=0)
    eng.set_option("pair_const", kset)
    eng.set_option("keep_hidden", 1)
    return eng


@pytest.mark.parametrize("name", ["decay_model_weights", "model2_weights", "model_weights"])
@pytest.mark.parametrize("dedup", [1, 0])
@pytest.mark.parametrize("kset", [0, 1, 2])
def test_pair_const_vs_oracle_and_default(engines, weights, mixed, name, dedup, kset):
    w = weights[name]
    rng = np.random.default_rng(44)
    idx = sorted(rng.choice(mixed.usable(w.n_x), 300, replace=False).tolist())
    offs, xyz, sp, Q = mixed.batch(idx, w.n_x)
    npad = np.where(np.arange(len(Q)) % 3 == 0, np.diff(offs), 41).astype(np.int32)      # some systems without padding
    eng = _engine(weights, name, kset)
    eng.set_option("dedup_far", dedup)
    try:
        q64 = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1].copy()
        h = eng.hidden(int(offs[-1])).copy()
        again = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1]
        assert np.array_equal(q64, again)
    finally:
        eng.close()
    ref = O.predict_batch(w, offs, xyz, sp, Q, npad)
    assert np.abs(q64 - ref).max() < TOL_FP32[name], (name, np.abs(q64 - ref).max())
    assert np.abs(np.add.reduceat(q64, offs[:-1]) - Q).max() < 1e-6
    simt = engines(name, 32)
    simt.set_option("keep_hidden", 1)
    q_simt = simt.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1]
    h_simt = simt.hidden(int(offs[-1]))
    assert np.abs(q64 - q_simt).max() < TOL_FP32[name]
    assert np.abs(h - h_simt).max() <= 2e-5 * max(1.0, np.abs(h_simt).max())


@pytest.mark.parametrize("name", ["model_weights", "model2_weights"])
@pytest.mark.parametrize("kset", [0, 1, 2])
def test_pair_const_hidden_state_vs_oracle(weights, mixed, name, kset):
    w = weights[name]
    idx = mixed.usable(w.n_x)[:40].tolist()
    offs, xyz, sp, Q = mixed.batch(idx, w.n_x)
    eng = _engine(weights, name, kset)
    try:
        eng.infer_batch(offs, xyz, sp, Q, 41)
        h = eng.hidden(int(offs[-1])).copy()
    finally:
        eng.close()
    for k in range(len(idx)):
        tr = {}
        a0, a1 = offs[k], offs[k + 1]
        O.forward_factorised(w, xyz[a0:a1], sp[a0:a1], Q[k], 41, trace=tr)
        assert np.abs(h[a0:a1] - tr["h"]).max() < 2e-4 * max(1.0, np.abs(tr["h"]).max()), (name, k)


@pytest.mark.parametrize("kset", [0, 1, 2])
def test_pair_const_golden_871(weights, mixed, val871, kset):
    w = weights["decay_model_weights"]
    idx = [mixed.index[n] for n in val871["names"]]
    offs, xyz, sp, Q = mixed.batch(idx, 9)
    eng = _engine(weights, "decay_model_weights", kset)
    try:
        q = eng.infer_batch(offs, xyz, sp, Q, 41)
    finally:
        eng.close()
    worst = max(float(np.abs(q[offs[k]:offs[k + 1]] - val871["pred"][k][:offs[k + 1] - offs[k]]).max()) for k in range(len(idx)))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("kset", [1, 2])
def test_pair_const_protein_golden(weights, protein, kset):
    """Galectin-3C through pair_const: per-atom variant, row-per-thread far kernel (epnn_gnn_far_const.cu) for the two
    live steps, species slots for the three collapsed ones."""
    eng = _engine(weights, "decay_model_weights", kset)
    try:
        n = len(protein["Z"])
        offs = np.array([0, n], np.int32)
        sp = O.species_from_Z(protein["Z"], 9)
        q, q64 = eng.infer_batch(offs, protein["xyz"], sp, np.array([protein["Q"]], np.float32), None, want_f64=True)
        assert eng.last_stats["n_far_dedup_rows"] == 3 * n
    finally:
        eng.close()
    assert np.abs(q - protein["preds"]).max() < 1e-5
    assert abs(q64.sum() - 2.0) < 1e-6


@pytest.mark.parametrize("name,rtol", [("model2_weights", 2e-4), ("model_weights", 2e-2)])
def test_pair_const_large_live_gnn(engines, weights, protein, name, rtol):
    w = weights[name]
    n = 600
    xyz = protein["xyz"][:n]
    sp = O.species_from_Z(protein["Z"][:n], w.n_x)
    offs = np.array([0, n], np.int32)
    Q = np.array([1.0], np.float32)
    eng = _engine(weights, name)
    try:
        for npad in (None, 640):
            q64 = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1].copy()
            h = eng.hidden(n).copy()
            tr = {}
            ref = O.forward_factorised(w, xyz, sp, Q[0], npad, trace=tr)
            assert np.abs(q64 - ref).max() / np.abs(ref).max() < rtol, (name, npad)
            assert np.abs(h - tr["h"]).max() / np.abs(tr["h"]).max() < 2e-4, (name, npad)
    finally:
        eng.close()
