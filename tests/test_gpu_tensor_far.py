"""Optional tensor-core path ("gnn_far_tensor" = 1): the far part of the big-system message sum on tcgen05 with the
3xTF32 split (epnn_gnn_tc.cu).  It must stay inside the FP32 path's tolerances against the float64 oracle, agree with
the FP32 SIMT path to the level of FP32 round-off, and be bitwise reproducible."""
import numpy as np
import pytest

from oracle import epnn_oracle as O

pytestmark = pytest.mark.gpu


def _engine(weights, name, impl=1):
    from epnn_b200.engine import Engine
    eng = Engine(weights[name], device=0)
    eng.set_option("gnn_far_tensor", 1)
    eng.set_option("gnn_far_tensor_impl", impl)      # 1: round-1 kernel (epnn_gnn_tc.cu), 2: warp-specialised pipeline (epnn_gnn_tc2.cu)
    eng.set_option("keep_hidden", 1)
    return eng


@pytest.mark.parametrize("impl", [1, 2])
def test_protein_golden_with_tensor_far(weights, protein, impl):
    eng = _engine(weights, "decay_model_weights", impl)
    n = len(protein["Z"])
    offs = np.array([0, n], np.int32)
    sp = O.species_from_Z(protein["Z"], 9)
    q, q64 = eng.infer_batch(offs, protein["xyz"], sp, np.array([protein["Q"]], np.float32), None, want_f64=True)
    assert np.abs(q - protein["preds"]).max() < 1e-5
    assert abs(q64.sum() - 2.0) < 1e-6
    eng.close()


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("name,rtol_q,rtol_h", [("model2_weights", 2e-4, 2e-5), ("model_weights", 2e-2, 2e-5)])
def test_live_gnn_tensor_far_vs_oracle_and_simt(weights, protein, engines, name, rtol_q, rtol_h, impl):
    w = weights[name]
    n = 700
    xyz = protein["xyz"][:n]
    sp = O.species_from_Z(protein["Z"][:n], w.n_x)
    offs = np.array([0, n], np.int32)
    Q = np.array([1.0], np.float32)
    eng = _engine(weights, name, impl)
    simt = engines(name)
    for npad in (None, 730):
        q64 = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1].copy()
        h_tc = eng.hidden(n).copy()
        again = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)[1]
        assert np.array_equal(q64, again)                                    # deterministic
        tr = {}
        ref = O.forward_factorised(w, xyz, sp, Q[0], npad, trace=tr)
        assert np.abs(q64 - ref).max() / np.abs(ref).max() < rtol_q, (name, npad)
        assert np.abs(h_tc - tr["h"]).max() / np.abs(tr["h"]).max() < 10 * rtol_h, (name, npad)
        simt.set_option("keep_hidden", 1)
        simt.infer_batch(offs, xyz, sp, Q, npad)
        h_simt = simt.hidden(n)
        # 3xTF32 vs FP32 SIMT: same order as the SIMT path's own distance to float64
        assert np.abs(h_tc - h_simt).max() / np.abs(h_simt).max() < rtol_h, (name, npad)
    eng.close()
