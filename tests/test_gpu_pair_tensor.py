"""Optional tensor variant of the electron-passing bundle kernel ("pair_tensor" = 1; epnn_bundle_mma.cu): both products
of the pair MLP on mma.sync TF32 with the 3xTF32 split, chained through registers.  It must stay inside the FP32 path's
tolerances against the float64 oracle, agree with the FP32 SIMT kernel to FP32 round-off, conserve charge, be bitwise
reproducible, and leave large systems (which take other kernels) untouched."""
import numpy as np
import pytest

from oracle import epnn_oracle as O

pytestmark = pytest.mark.gpu

TOL_FP32 = {"decay_model_weights": 3e-6, "model2_weights": 1e-5, "model_weights": 2e-4}     # FP32 floors of tests/test_gpu_parity.py (+ the 3xTF32 split for decay)


@pytest.mark.parametrize("name", ["decay_model_weights", "model2_weights", "model_weights"])
def test_pair_tensor_vs_oracle_and_simt(engines, weights, mixed, name):
    from epnn_b200.engine import Engine
    w = weights[name]
    rng = np.random.default_rng(33)
    idx = sorted(rng.choice(mixed.usable(w.n_x), 300, replace=False).tolist())
    offs, xyz, sp, Q = mixed.batch(idx, w.n_x)
    eng = Engine(w, device=0)
    eng.set_option("pair_tensor", 1)
    try:
        q, q64 = eng.infer_batch(offs, xyz, sp, Q, 41, want_f64=True)
        q64 = q64.copy()
        again = eng.infer_batch(offs, xyz, sp, Q, 41, want_f64=True)[1]
        assert np.array_equal(q64, again)                                   # deterministic
    finally:
        eng.close()
    ref = O.predict_batch(w, offs, xyz, sp, Q, np.full(len(Q), 41))
    assert np.abs(q64 - ref).max() < TOL_FP32[name], (name, np.abs(q64 - ref).max())
    sums = np.add.reduceat(q64, offs[:-1])
    assert np.abs(sums - Q).max() < 1e-6                                    # antisymmetric transfers: conserved by construction
    simt = engines(name, 32).infer_batch(offs, xyz, sp, Q, 41, want_f64=True)[1]
    assert 0 < np.abs(q64 - simt).max() < TOL_FP32[name], (name, np.abs(q64 - simt).max())   # another kernel really ran


def test_pair_tensor_golden_871(weights, mixed, val871):
    """The reference's own shipped predictions (decay_model_weights, pad 41) through the tensor variant."""
    from epnn_b200.engine import Engine
    w = weights["decay_model_weights"]
    idx = [mixed.index[n] for n in val871["names"]]
    offs, xyz, sp, Q = mixed.batch(idx, 9)
    eng = Engine(w, device=0)
    eng.set_option("pair_tensor", 1)
    try:
        q = eng.infer_batch(offs, xyz, sp, Q, 41)
    finally:
        eng.close()
    worst = 0.0
    for k in range(len(idx)):
        n = offs[k + 1] - offs[k]
        worst = max(worst, float(np.abs(q[offs[k]:offs[k + 1]] - val871["pred"][k][:n]).max()))
    assert worst < 1e-5, worst
