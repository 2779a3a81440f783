"""CPU oracle: a numpy restatement of the reference's EPNN charge-inference path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``epnn_b200``) never imports this module and has no CPU fallback.

Parity status:
  * pinned for ``decay_model_weights`` by the reference's own shipped predictions:
    871 validation systems (``models/model_systems/test_pred_charges.npy``, pad N=41) and the
    2220-atom Galectin-3C protein (``data/protein.tar.gz -> protein/preds.npy``); see
    ``tests/test_oracle_golden.py``.
  * ``model_weights`` / ``model2_weights``: PARITY UNPINNED (no shipped array matches them), and the
    GNN-layer arithmetic is unpinned for every checkpoint because ``decay_model_weights``' GNN output
    is a dead constant.  For those the oracle follows the reference source literally (citations below).

TensorFlow is not installable in this image, so the reference itself cannot be executed; every
function cites the reference ``file:line`` it restates.

Two formulations are provided:
  * ``forward_literal``     -- the dense, padded, un-factorised graph exactly as ``charge_gn.py`` builds
                               it (explicit (N*N, K) pair inputs, three Dense layers per MLP).
  * ``forward_factorised``  -- an algebraically identical rewrite that is fast enough for
                               the 2220-atom protein; tests assert it equals ``forward_literal`` to 1e-11.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

E_DIM = 48
# Element tables restated from the reference so the oracle is independent of the product package:
# charge_gn.py:9-28 (10-wide features, with P) and infer.py:13-30 (9-wide, without P).
_ATOMIC_NUMBER = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'P': 15, 'S': 16, 'Cl': 17, 'Br': 35}
_TABLES = {10: ('H', 'C', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br'), 9: ('H', 'C', 'N', 'O', 'F', 'S', 'Cl', 'Br')}


def features(species, n_x):
    """x_i = [Z, onehot(species)] (charge_gn.py:325-328); ``species`` indexes the table chosen by n_x."""
    species = np.asarray(species)
    z = np.array([_ATOMIC_NUMBER[s] for s in _TABLES[n_x]], dtype=np.float32)
    x = np.zeros((species.shape[0], n_x), dtype=np.float32)
    x[:, 0] = z[species]
    x[np.arange(species.shape[0]), species + 1] = 1.0
    return x


def species_from_symbols(symbols, n_x):
    table = {s: i for i, s in enumerate(_TABLES[n_x])}
    return np.array([table[s] for s in symbols], dtype=np.int32)


def species_from_Z(Z, n_x):
    zt = [_ATOMIC_NUMBER[s] for s in _TABLES[n_x]]
    return np.array([zt.index(int(z)) for z in Z], dtype=np.int32)


CUTOFF = 3.0
ETA = 2.0
NEAR_TOL = np.float32(1e-5)       # tf.constant(1e-5) is float32: charge_gn.py:90


# ------------------------------------------------------------------------------------ descriptors
def distance_matrix(xyz: np.ndarray) -> np.ndarray:
    """``scipy.spatial.distance_matrix(xyz, xyz)`` (charge_gn.py:124) restated.

    scipy promotes float32 input to float64, takes |y - x|**2, sums the three components in order
    ((dx2 + dy2) + dz2) and applies ``**0.5`` (numpy evaluates that as sqrt)."""
    x = np.asarray(xyz).astype(np.float64)
    d = np.abs(x[None, :, :] - x[:, None, :])
    sq = d * d
    s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    return np.sqrt(s)


def get_init_edges(xyz: np.ndarray, num: int = E_DIM, cutoff: float = CUTOFF, eta: float = ETA):
    """charge_gn.py:122-163.  Returns (e (n,n,num) float32, C (n,n) float64).

    The ``adj`` / molecular_splits block (:126-146) is dead code upstream and is not restated."""
    mu = np.linspace(0.1, cutoff, num=num)                       # :123
    D = distance_matrix(xyz)                                      # :124
    C = (np.cos(np.pi * (D - 0.0) / cutoff) + 1.0) / 2.0          # :148
    C[D >= cutoff] = 0.0                                          # :150
    C[D <= 0.0] = 1.0                                             # :151
    np.fill_diagonal(C, 0.0)                                      # :152
    e = C[..., None] * np.exp(-eta * (D[..., None] - mu[None, None, :]) ** 2)   # :153-160
    return np.array(e, dtype=np.float32), C                       # :161


def is_near_from_e(e: np.ndarray) -> np.ndarray:
    """charge_gn.py:90-94: clip(e,1e-5,1e5).max(-1) != 1e-5  <=>  max_k e_ijk > float32(1e-5)."""
    e32 = np.asarray(e, dtype=np.float32)
    clip = np.clip(e32, NEAR_TOL, np.float32(1e5))
    return clip.max(axis=-1) != NEAR_TOL


def neighbor_csr(xyz: np.ndarray):
    """Sorted CSR of the ``is_near`` set for one system (the list the CUDA build must match bit-exactly)."""
    e, _ = get_init_edges(xyz)
    near = is_near_from_e(e)
    n = near.shape[0]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(near.sum(axis=1))
    col = np.nonzero(near)[1].astype(np.int32)
    return rowptr, col


def initial_charge(Q, n: int) -> np.float32:
    """charge_gn.py:317,337-338: q0 = float32(float32(Q) / n) on every real atom."""
    return np.float32(np.float32(Q) / np.float32(n))


# ------------------------------------------------------------------------------------ layers
def _relu(x):
    return np.maximum(x, 0)


def _mlp(mlp, x, dt):
    """MLP_layer.call, charge_gn.py:41-45: Dense(relu)... then Dense(None); y = x @ kernel + bias."""
    n = len(mlp.W)
    for i, (W, b) in enumerate(zip(mlp.W, mlp.b)):
        x = x @ W.astype(dt) + b.astype(dt)
        if i < n - 1:
            x = _relu(x)
    return x


def _pair_rows(a, e_rows, r0, r1, swap=False):
    """Rows r0:r1 of the (N,N,K) pair-input tensor [a_i | a_j | e_ij] (charge_gn.py:62-66 / 101-108)."""
    N = a.shape[0]
    R = r1 - r0
    ai = np.broadcast_to(a[r0:r1, None, :], (R, N, a.shape[1]))
    aj = np.broadcast_to(a[None, :, :], (R, N, a.shape[1]))
    if swap:
        ai, aj = aj, ai
    return np.concatenate([ai, aj, e_rows], axis=-1).reshape(R * N, -1)


def _chunk(N, K, budget=40_000_000):
    return max(1, min(N, budget // max(1, N * K)))


def gnn_layer_literal(w, h, e, x, q, mask, dt, trace=None):
    """GNN_layer.call, charge_gn.py:56-75 (single system; leading batch dim dropped)."""
    N = e.shape[0]
    node_mask = np.clip(mask.sum(axis=0), 0, 1).astype(dt)[:, None]                 # :59
    rc = _chunk(N, w.K)
    for t in range(w.T):                                                            # :60
        a = np.concatenate([x, h, q], axis=-1)                                      # :62
        messages = np.zeros((N, 32), dtype=dt)
        for r0 in range(0, N, rc):
            r1 = min(N, r0 + rc)
            inp = _pair_rows(a, e[r0:r1].astype(dt), r0, r1)                        # :63-66
            m = _mlp(w.msg[t], inp, dt).reshape(r1 - r0, N, 32)                     # :68-69
            messages[r0:r1] = m.sum(axis=1)                                         # :70  (NO mask)
        upd_in = np.concatenate([h, messages], axis=1) * node_mask                 # :71-72
        h = _mlp(w.upd, upd_in, dt) * node_mask                                     # :73-74
        if trace is not None:
            trace.setdefault("messages", []).append(messages.copy())
            trace.setdefault("h_steps", []).append(h.copy())
    return h


def epn_layer_literal(w, h, e, x, q, mask, dt, trace=None):
    """EPN_layer.call, charge_gn.py:87-119."""
    N = e.shape[0]
    near = is_near_from_e(e).astype(dt)                                             # :90-94
    maskf = mask.astype(dt)                                                         # :97
    rc = _chunk(N, w.K)
    q = q.copy()
    for t in range(w.T):                                                            # :98
        a = np.concatenate([x, h, q], axis=-1)                                      # :101
        dq = np.zeros((N,), dtype=dt)
        for r0 in range(0, N, rc):
            r1 = min(N, r0 + rc)
            er = e[r0:r1].astype(dt)
            f_ij = _mlp(w.pas[t], _pair_rows(a, er, r0, r1), dt).reshape(r1 - r0, N)              # :104,107,110
            f_ji = _mlp(w.pas[t], _pair_rows(a, er, r0, r1, swap=True), dt).reshape(r1 - r0, N)   # :105,108,111
            anti = 0.5 * (f_ij - f_ji) * maskf[r0:r1] * near[r0:r1]                 # :116
            dq[r0:r1] = anti.sum(axis=1)
        q = q + dq[:, None]                                                         # :118
        if trace is not None:
            trace.setdefault("q_steps", []).append(q[:, 0].copy())
    return q


def model_forward_keras_inputs(w, h_inp, e_inp, x_inp, q_inp, mask, dtype=np.float64, trace=None):
    """make_model wiring, charge_gn.py:382-387, for ONE system of Keras-shaped inputs:
    h_inp (N,N,48), e_inp (N,N,48), x_inp (N,N,n_x), q_inp (N,N,1), mask (N,N).  Returns q (N,1)."""
    dt = dtype
    mask = np.asarray(mask).reshape(mask.shape[0], mask.shape[1])
    den = mask.sum(axis=0).astype(dt)[:, None]                                      # reduce_sum(mask, axis=1)

    def untile(t):                                                                  # divide_no_nan, :382-384
        s = np.asarray(t).astype(dt).sum(axis=0)
        out = np.zeros_like(s)
        np.divide(s, den, out=out, where=den != 0)
        return out

    h, x, q = untile(h_inp), untile(x_inp), untile(q_inp)
    e = np.asarray(e_inp, dtype=np.float32)
    hg = gnn_layer_literal(w, h, e, x, q, mask, dt, trace)                          # :386
    if trace is not None:
        trace["h"] = hg.copy()
    return epn_layer_literal(w, hg, e, x, q, mask, dt, trace)                       # :387


# ------------------------------------------------------------------------------------ per-system drivers
def _padded_inputs(w, xyz, species, Q, npad, dt):
    """What gen_padded_init_state (charge_gn.py:292-366) + the un-tiling (:382-384) hand the layers for
    one system padded to ``npad``: per-atom x (N,n_x), h (N,48)=0, q (N,1)=q0, e (N,N,48) f32, mask (N,N)."""
    n = xyz.shape[0]
    N = int(npad)
    if N < n:
        raise ValueError("npad must be >= number of atoms")
    x = np.zeros((N, w.n_x), dtype=dt)
    x[:n] = features(species, w.n_x)
    h = np.zeros((N, w.h_dim), dtype=dt)
    q = np.zeros((N, 1), dtype=dt)
    q[:n, 0] = initial_charge(Q, n)
    e = np.zeros((N, N, E_DIM), dtype=np.float32)
    e[:n, :n] = get_init_edges(np.asarray(xyz, dtype=np.float32))[0]
    mask = np.zeros((N, N), dtype=dt)
    mask[:n, :n] = 1
    return x, h, q, e, mask


def forward_literal(w, xyz, species, Q, npad=None, dtype=np.float64, trace: Optional[Dict] = None):
    """Reference formulation end to end for one system; returns q for the n real atoms."""
    n = xyz.shape[0]
    npad = n if npad is None else npad
    x, h, q, e, mask = _padded_inputs(w, xyz, species, Q, npad, dtype)
    hg = gnn_layer_literal(w, h, e, x, q, mask, dtype, trace)
    if trace is not None:
        trace["h"] = hg[:n].copy()
    qf = epn_layer_literal(w, hg, e, x, q, mask, dtype, trace)
    return qf[:n, 0]


def forward_factorised(w, xyz, species, Q, npad=None, dtype=np.float64, trace: Optional[Dict] = None):
    """Same mathematics as ``forward_literal`` with the exact rewrites of SURVEY.md 7.2:
    first Dense layer split into per-atom projections (W1 = [A;B;C] row blocks), the linear last
    message layer hoisted out of the sum over j, padded atoms folded into one weighted pseudo-pair
    (a_j = 0, e = 0), and the EPN evaluated on unordered near pairs with +/- scatter."""
    dt = dtype
    xyz = np.asarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    N = n if npad is None else int(npad)
    x = features(species, w.n_x).astype(dt)
    h = np.zeros((n, w.h_dim), dtype=dt)
    q = np.full((n, 1), initial_charge(Q, n), dtype=dt)
    e32, _ = get_init_edges(xyz)
    F = w.F
    nz_i, nz_j = np.nonzero((e32 != 0).any(axis=-1))
    e_nz = e32[nz_i, nz_j].astype(dt)
    rc = max(1, min(n, 30_000_000 // max(1, n * 32)))
    for t in range(w.T):
        W1, b1 = w.msg[t].W[0].astype(dt), w.msg[t].b[0].astype(dt)
        W2, b2 = w.msg[t].W[1].astype(dt), w.msg[t].b[1].astype(dt)
        W3, b3 = w.msg[t].W[2].astype(dt), w.msg[t].b[2].astype(dt)
        a = np.concatenate([x, h, q], axis=-1)
        u = a @ W1[:F]
        v = a @ W1[F:2 * F] + b1
        ce = e_nz @ W1[2 * F:]
        S = np.zeros((n, 32), dtype=dt)
        for r0 in range(0, n, rc):
            r1 = min(n, r0 + rc)
            z = u[r0:r1, None, :] + v[None, :, :]
            sel = (nz_i >= r0) & (nz_i < r1)
            z[nz_i[sel] - r0, nz_j[sel]] += ce[sel]
            S[r0:r1] = _relu(_relu(z) @ W2 + b2).sum(axis=1)
        S += (N - n) * _relu(_relu(u + b1) @ W2 + b2)
        M = S @ W3 + N * b3
        h = _mlp(w.upd, np.concatenate([h, M], axis=1), dt)
        if trace is not None:
            trace.setdefault("messages", []).append(M.copy())
            trace.setdefault("h_steps", []).append(h.copy())
    if trace is not None:
        trace["h"] = h.copy()
    near = is_near_from_e(e32)
    pi, pj = np.nonzero(np.triu(near, 1))
    ep = e32[pi, pj].astype(dt)
    for t in range(w.T):
        W1, b1 = w.pas[t].W[0].astype(dt), w.pas[t].b[0].astype(dt)
        W2, b2 = w.pas[t].W[1].astype(dt), w.pas[t].b[1].astype(dt)
        W3 = w.pas[t].W[2].astype(dt)
        a = np.concatenate([x, h, q], axis=-1)
        u = a @ W1[:F]
        v = a @ W1[F:2 * F] + b1
        ce = ep @ W1[2 * F:]
        f_ij = _relu(_relu(u[pi] + v[pj] + ce) @ W2 + b2) @ W3
        f_ji = _relu(_relu(u[pj] + v[pi] + ce) @ W2 + b2) @ W3
        d = 0.5 * (f_ij - f_ji)[:, 0]
        dq = np.zeros(n, dtype=dt)
        np.add.at(dq, pi, d)
        np.add.at(dq, pj, -d)
        q = q + dq[:, None]
        if trace is not None:
            trace.setdefault("q_steps", []).append(q[:, 0].copy())
    return q[:, 0]


def predict_batch(w, offsets, xyz, species, Q, npad, dtype=np.float64, literal=False):
    """Run the oracle over a packed batch (same argument layout as the C-ABI ``epnn_infer_batch``)."""
    out = np.zeros(int(offsets[-1]), dtype=np.float64)
    fn = forward_literal if literal else forward_factorised
    for s in range(len(offsets) - 1):
        a0, a1 = int(offsets[s]), int(offsets[s + 1])
        out[a0:a1] = fn(w, xyz[a0:a1], species[a0:a1], Q[s], int(npad[s]), dtype)
    return out
