#!/bin/bash
# A/B of an engine option at NMOL (default 300 k) molecules:  gpurun -- '[CKPT=model2_weights] bash tools/gpu_ab_opt.sh atom_tensor 0 1'
mkdir -p gpurun_out
OPT=$1; shift
cat > /tmp/ab_opt.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth
opt, vals = sys.argv[1], [float(x) for x in sys.argv[2:]]
w = load_weights("tests/golden/checkpoints/" + os.environ.get("CKPT", "decay_model_weights"))
NMOL = int(os.environ.get("NMOL", "300000"))
offs, xyz, sp, Q = synth.qm9_shaped(NMOL, w.n_x, seed=0)
npad = np.full(NMOL, 29, np.int32)
outs = []
for v in vals:
    eng = Engine(w, 0); eng.set_option("timing", 1); eng.set_option(opt, v)
    acc = {}
    for it in range(6):
        q, q64 = eng.infer_batch(offs, xyz, sp, Q, npad, want_f64=True)
        if it >= 3:
            for k, x in eng.last_stats.items(): acc[k] = acc.get(k, 0) + x / 3
    print(opt, v, {k: round(x, 2) for k, x in acc.items() if k.startswith("ms_")}, flush=True)
    outs.append(q64.copy()); eng.close()
print("bit-identical across values:", all(np.array_equal(outs[0], o) for o in outs))
PY
timeout 300 python /tmp/ab_opt.py $OPT "$@"
