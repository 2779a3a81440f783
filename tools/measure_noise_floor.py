"""Measured distance of every precision mode to the float64 oracle (development aid; the output is the evidence behind the
tolerances in tests/test_gpu_parity.py and DESIGN.md section 2).   python tools/measure_noise_floor.py [n_systems] [OPTION=VALUE ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Mixed, GOLDEN, CKPTS
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from oracle import epnn_oracle as O

n_sys = int(sys.argv[1]) if len(sys.argv) > 1 else 400
opts = [a.split("=") for a in sys.argv[2:]]
mx = Mixed()
for name in CKPTS:
    w = load_weights(os.path.join(GOLDEN, "checkpoints", name))
    rng = np.random.default_rng(11)
    idx = sorted(rng.choice(mx.usable(w.n_x), n_sys, replace=False).tolist())
    offs, xyz, sp, Q = mx.batch(idx, w.n_x)
    for npad_mode in ("41", "n"):
        npads = np.full(len(idx), 41, np.int32) if npad_mode == "41" else np.diff(offs).astype(np.int32)
        ref = O.predict_batch(w, offs, xyz, sp, Q, npads)
        for prec in (32, 48, 64, 0):
            eng = Engine(w, device=0, precision=prec)
            eng.set_option("timing", 1)
            for k, v in opts:
                eng.set_option(k, float(v))
            q, q64 = eng.infer_batch(offs, xyz, sp, Q, npads, want_f64=True)
            st = eng.last_stats
            per_sys = np.maximum.reduceat(np.abs(q64 - ref), offs[:-1])
            print(f"{name:20s} pad {npad_mode:2s} precision {prec:2d} -> used {st['precision_used']:2d}  max|dq| {per_sys.max():.2e}  p99 {np.quantile(per_sys, 0.99):.2e}  "
                  f"median {np.median(per_sys):.2e}  |sum q - Q| {np.abs(np.add.reduceat(q64, offs[:-1]) - Q).max():.1e}  probe32 {st['probe_err32']:.1e} probe48 {st['probe_err48']:.1e}  ms {st['ms_total']:.2f}", flush=True)
            eng.close()
