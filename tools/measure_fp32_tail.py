"""Tail of the FP32 error distribution over ALL usable systems of data/mixed (development aid): max |dq| per system of the FP32
kernels (per-atom kernel on the tensor path / SIMT) and of the mixed precision against the FP64 kernels (which agree with the
float64 oracle to 1e-9, tests/test_gpu_parity.py).   python tools/measure_fp32_tail.py [checkpoint ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Mixed, GOLDEN, CKPTS
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine

mx = Mixed()
for name in (sys.argv[1:] or CKPTS):
    w = load_weights(os.path.join(GOLDEN, "checkpoints", name))
    idx = mx.usable(w.n_x).tolist()
    offs, xyz, sp, Q = mx.batch(idx, w.n_x)
    for pad in ("41", "n"):
        npads = np.full(len(idx), 41, np.int32) if pad == "41" else np.diff(offs).astype(np.int32)
        res = {}
        for label, prec, opts in (("fp64", 64, {}), ("fp32 tensor", 32, {"atom_tensor": 1}), ("fp32 simt", 32, {"atom_tensor": 0}), ("mixed", 48, {})):
            eng = Engine(w, device=0, precision=prec)
            for k, v in opts.items():
                eng.set_option(k, v)
            res[label] = eng.infer_batch(offs, xyz, sp, Q, npads, want_f64=True)[1].copy()
            eng.close()
        for label in ("fp32 tensor", "fp32 simt", "mixed"):
            per = np.maximum.reduceat(np.abs(res[label] - res["fp64"]), offs[:-1])
            print(f"{name:20s} pad {pad:2s} {label:12s} n={len(idx)}  max {per.max():.2e}  p99.9 {np.quantile(per, 0.999):.2e}  p99 {np.quantile(per, 0.99):.2e}  "
                  f"median {np.median(per):.2e}  >1e-5: {(per > 1e-5).sum()}  >5e-6: {(per > 5e-6).sum()}  >2.5e-6: {(per > 2.5e-6).sum()}", flush=True)
