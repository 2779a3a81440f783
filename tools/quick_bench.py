"""Quick per-phase timing of the batched small-molecule path (development aid, not the bench contract)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ck = sys.argv[2] if len(sys.argv) > 2 else "decay_model_weights"
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 32
w = load_weights(os.path.join(ROOT, "tests", "golden", "checkpoints", ck))
t0 = time.time()
offs, xyz, sp, Q = synth.qm9_shaped(n_mol, w.n_x, seed=0)
print(f"generated {n_mol} molecules / {offs[-1]} atoms in {time.time()-t0:.2f}s")
eng = Engine(w, 0, prec)
eng.set_option("timing", 1)
npad = np.full(n_mol, 29, np.int32)
for it in range(4):
    t0 = time.time()
    q = eng.infer_batch(offs, xyz, sp, Q, npad)
    dt = time.time() - t0
    st = eng.last_stats
    print(f"iter {it}: wall {dt*1e3:.1f} ms  atoms/s {offs[-1]/dt:.3e}  " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if k.startswith("ms_")))
print({k: v for k, v in st.items() if not k.startswith("ms_")})
print("sum|q| per mol max", np.abs(np.add.reduceat(q.astype(np.float64), offs[:-1])).max())
if len(sys.argv) > 4:
    x, Z, Qp = synth.protein()
    spp = synth.species_from_Z(Z, w.n_x)
    o = np.array([0, len(Z)], np.int32)
    for it in range(3):
        t0 = time.time(); q = eng.infer_batch(o, x, spp, np.array([Qp], np.float32)); dt = time.time() - t0
        st = eng.last_stats
        print(f"protein iter {it}: wall {dt*1e3:.1f} ms " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if k.startswith("ms_")))
