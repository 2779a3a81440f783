"""Small end-to-end case for compute-sanitizer (development aid): bundles, a brute-force large system, the cell-list
path, the dense compat path and the tcgen05 far kernel, each once, at sizes that finish in seconds under memcheck."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth

w = load_weights(os.path.join(ROOT, "tests", "golden", "checkpoints", "model2_weights"))
eng = Engine(w, 0)
offs, xyz, sp, Q = synth.qm9_shaped(40, 9, seed=1)
px, pZ, pQ = synth.protein()
psp = synth.species_from_Z(pZ, 9)
n1, n2 = 130, 600
offs2 = np.concatenate([offs, [offs[-1] + n1, offs[-1] + n1 + n2]]).astype(np.int32)
xyz2 = np.concatenate([xyz, px[:n1], px[200:200 + n2]]).astype(np.float32)
sp2 = np.concatenate([sp, psp[:n1], psp[200:200 + n2]]).astype(np.int32)
Q2 = np.concatenate([Q, [1.0, -1.0]]).astype(np.float32)
npad = np.concatenate([np.full(40, 29), [n1 + 3, n2]]).astype(np.int32)
a = eng.infer_batch(offs2, xyz2, sp2, Q2, npad)
eng.set_option("gnn_far_tensor", 1)
b = eng.infer_batch(offs2, xyz2, sp2, Q2, npad)
print("simt vs tensor max diff", float(np.abs(a - b).max()))
rp, col = eng.neighbors(offs2, xyz2, 0)
e = eng.init_edges(xyz2[:20])
N = 12
rng = np.random.default_rng(0)
out = eng.infer_dense(rng.normal(size=(1, N, N, 48)), np.abs(rng.normal(size=(1, N, N, 48))), rng.normal(size=(1, N, N, 9)),
                      rng.normal(size=(1, N, N)), (rng.random((1, N, N)) > 0.3).astype(np.float32))
print("ok", a.shape, col.shape, e.shape, out.shape)
