"""Pure-Python reader (and template-based writer) for the TensorFlow "tensor bundle" checkpoints the reference ships.

Replaces ``model.load_weights('./models/decay_model_weights')`` (reference ``infer.py:57``) and
``model.save_weights(...)`` (reference ``charge_gn.py:462``) without TensorFlow.  Format:

* ``<prefix>.index`` is a LevelDB-style SSTable: 48-byte footer (metaindex handle, index handle,
  magic 0xdb4775248b80fb57), prefix-compressed key/value blocks, each followed by a 1-byte compression
  tag and a 4-byte masked CRC32C.
* key ``""`` holds a ``BundleHeaderProto`` (field 1 = num_shards); every other key holds a
  ``BundleEntryProto`` (1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 crc32c).
* tensor bytes are raw little-endian, row-major, in ``<prefix>.data-%05d-of-%05d``.

Variable naming (Keras object graph of ``charge_gn.make_model``, reference ``charge_gn.py:369-391``):
``layer_with_weights-0`` = ``GNN_layer`` (``message_fns/<t>``, ``message_fn``, ``update_fn``),
``layer_with_weights-1`` = ``EPN_layer`` (``pass_fns/<t>``, ``pass_fn``).  Because ``call`` rebinds
``self.message_fn = self.message_fns[t]`` (``charge_gn.py:61``, ``:99``) the LAST step's MLP is stored
under the un-indexed attribute, steps ``0..T-2`` under the indexed list.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


class CheckpointError(ValueError):
    """Raised for malformed / unsupported checkpoint files."""


# ----------------------------------------------------------------------------- low level: varints
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


# ----------------------------------------------------------------------------- CRC32C (Castagnoli)
def _make_crc_table() -> List[int]:
    poly = 0x82F63B78
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- minimal protobuf wire reader
def _proto_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """Decode one protobuf message into [(field_number, wire_type, value)] (no schema)."""
    out = []
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fnum, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            if len(val) != ln:
                raise CheckpointError("truncated length-delimited field")
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wt}")
        out.append((fnum, wt, val))
    return out


# ----------------------------------------------------------------------------- SSTable
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise CheckpointError("block handle points outside the index file")
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise CheckpointError(f"compressed index blocks (type {ctype}) are not supported")
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if stored != masked_crc32c(data[offset:offset + size + 1]):
            raise CheckpointError("index block CRC mismatch")
    return block


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise CheckpointError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos = 0
    key = b""
    out = []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(path: str, verify_crc: bool = True) -> Dict[str, bytes]:
    """Return {key: raw value bytes} for every entry of an ``.index`` SSTable."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise CheckpointError("index file shorter than its footer")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise CheckpointError("bad SSTable magic; not a TF checkpoint index")
    pos = 0
    _, pos = _varint(footer, pos)      # metaindex offset
    _, pos = _varint(footer, pos)      # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, bytes] = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_crc)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for k, v in _block_entries(_read_block(data, boff, bsize, verify_crc)):
            entries[k.decode("utf-8")] = v
    return entries


# ----------------------------------------------------------------------------- bundle entries
@dataclass
class BundleEntry:
    dtype: int = 0
    shape: Tuple[int, ...] = ()
    shard_id: int = 0
    offset: int = 0
    size: int = 0
    crc32c: int = 0


def _parse_entry(raw: bytes) -> BundleEntry:
    e = BundleEntry()
    for fnum, _, val in _proto_fields(raw):
        if fnum == 1:
            e.dtype = val
        elif fnum == 2:
            dims = []
            for f2, _, v2 in _proto_fields(val):
                if f2 == 2:      # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            e.shape = tuple(dims)
        elif fnum == 3:
            e.shard_id = val
        elif fnum == 4:
            e.offset = val
        elif fnum == 5:
            e.size = val
        elif fnum == 6:
            e.crc32c = val
    return e


def read_bundle(prefix: str, verify_crc: bool = True) -> Dict[str, np.ndarray]:
    """Read every DT_FLOAT tensor of the checkpoint ``prefix`` → {variable key: float32 ndarray}."""
    index = read_index(prefix + ".index", verify_crc)
    if "" not in index:
        raise CheckpointError("bundle header entry missing")
    num_shards = 1
    for fnum, _, val in _proto_fields(index[""]):
        if fnum == 1:
            num_shards = val
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for key, raw in index.items():
        if key == "":
            continue
        ent = _parse_entry(raw)
        if ent.dtype != 1:          # DT_FLOAT only; the object graph (DT_STRING) is skipped
            continue
        if ent.shard_id not in shards:
            p = f"{prefix}.data-{ent.shard_id:05d}-of-{num_shards:05d}"
            if not os.path.exists(p):
                raise CheckpointError(f"missing shard file {p}")
            with open(p, "rb") as f:
                shards[ent.shard_id] = f.read()
        blob = shards[ent.shard_id][ent.offset:ent.offset + ent.size]
        n = int(np.prod(ent.shape)) if ent.shape else 1
        if len(blob) != ent.size or ent.size != 4 * n:
            raise CheckpointError(f"tensor {key}: size/shape mismatch")
        if verify_crc and ent.crc32c and masked_crc32c(blob) != ent.crc32c:
            raise CheckpointError(f"tensor {key}: data CRC mismatch")
        out[key] = np.frombuffer(blob, dtype="<f4").reshape(ent.shape).copy()
    return out


# ----------------------------------------------------------------------------- model-level view
@dataclass
class MLP:
    """One ``MLP_layer`` (reference ``charge_gn.py:30-45``): kernels are (in, out), y = x @ W + b."""
    W: List[np.ndarray] = field(default_factory=list)
    b: List[np.ndarray] = field(default_factory=list)


@dataclass
class Weights:
    T: int
    n_x: int
    h_dim: int
    e_dim: int
    msg: List[MLP]       # T message MLPs  K -> 32 -> 32 -> 32
    upd: MLP             # shared update MLP 80 -> 32 -> 32 -> 48
    pas: List[MLP]       # T pass MLPs     K -> 32 -> 32 -> 1

    @property
    def F(self) -> int:           # per-atom input width [x | h | q]
        return self.n_x + self.h_dim + 1

    @property
    def K(self) -> int:           # pair input width [a_i | a_j | e_ij]
        return 2 * self.F + self.e_dim

    def packed(self) -> np.ndarray:
        """Flat float32 buffer in the order ``epnn_create`` expects (include/epnn_b200.h)."""
        parts = []
        for m in self.msg:
            for W, b in zip(m.W, m.b):
                parts += [W.ravel(), b.ravel()]
        for W, b in zip(self.upd.W, self.upd.b):
            parts += [W.ravel(), b.ravel()]
        for m in self.pas:
            for W, b in zip(m.W, m.b):
                parts += [W.ravel(), b.ravel()]
        return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def _collect_mlp(tensors: Dict[str, np.ndarray], base: str) -> MLP:
    mlp = MLP()
    i = 0
    while f"{base}/layer_set/{i}/kernel{_SUFFIX}" in tensors:
        mlp.W.append(tensors[f"{base}/layer_set/{i}/kernel{_SUFFIX}"])
        mlp.b.append(tensors[f"{base}/layer_set/{i}/bias{_SUFFIX}"])
        i += 1
    if i == 0:
        raise CheckpointError(f"no layers found under {base}")
    return mlp


def load_weights(prefix: str, verify_crc: bool = True) -> Weights:
    """Load a reference checkpoint and resolve the step→variable aliasing (see module docstring)."""
    tensors = read_bundle(prefix, verify_crc)
    g, p = "layer_with_weights-0", "layer_with_weights-1"
    steps = sorted({int(m.group(1)) for k in tensors
                    for m in [re.match(rf"{g}/message_fns/(\d+)/", k)] if m})
    T = len(steps) + 1
    if steps != list(range(T - 1)):
        raise CheckpointError(f"non-contiguous message_fns indices {steps}")
    msg = [_collect_mlp(tensors, f"{g}/message_fns/{t}") for t in range(T - 1)]
    msg.append(_collect_mlp(tensors, f"{g}/message_fn"))
    pas = [_collect_mlp(tensors, f"{p}/pass_fns/{t}") for t in range(T - 1)]
    pas.append(_collect_mlp(tensors, f"{p}/pass_fn"))
    upd = _collect_mlp(tensors, f"{g}/update_fn")
    h_dim = upd.W[-1].shape[1]
    K = msg[0].W[0].shape[0]
    e_dim = 48                                   # get_init_edges(num=e_dim), reference charge_gn.py:331
    F2 = K - e_dim
    if F2 % 2:
        raise CheckpointError(f"unexpected first-layer width {K}")
    n_x = F2 // 2 - h_dim - 1
    w = Weights(T=T, n_x=n_x, h_dim=h_dim, e_dim=e_dim, msg=msg, upd=upd, pas=pas)
    for m in msg:
        if [x.shape for x in m.W] != [(K, 32), (32, 32), (32, 32)]:
            raise CheckpointError("unexpected message MLP shapes")
    for m in pas:
        if [x.shape for x in m.W] != [(K, 32), (32, 32), (32, 1)]:
            raise CheckpointError("unexpected pass MLP shapes")
    if [x.shape for x in upd.W] != [(h_dim + 32, 32), (32, 32), (32, h_dim)]:
        raise CheckpointError("unexpected update MLP shapes")
    return w


# ----------------------------------------------------------------------------- writer (template based)
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    """LevelDB block: prefix-compressed entries, a restart point every ``restart_interval`` entries."""
    buf = bytearray()
    restarts = []
    last = b""
    for n, (key, val) in enumerate(entries):
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(buf))
        else:
            m = min(len(last), len(key))
            while shared < m and last[shared] == key[shared]:
                shared += 1
        buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def _short_successor(key: bytes) -> bytes:
    """BytewiseComparator::FindShortSuccessor: first byte that is not 0xff incremented, the rest dropped."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _write_index(path: str, entries: List[Tuple[bytes, bytes]]):
    """One-data-block SSTable exactly as TensorFlow's TableBuilder lays it out for a bundle this small:
    data block, empty metaindex block, index block with one entry (short successor of the last key), 48-byte footer."""
    out = bytearray()

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(out)
        out.extend(block)
        out.append(0)                                                   # no compression
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return off, len(block)

    data = _build_block(entries)
    if len(data) > 256 * 1024:
        raise CheckpointError("index would need more than one data block (not produced by the reference's models)")
    d_off, d_size = emit(data)
    m_off, m_size = emit(_build_block([]))
    handle = _put_varint(d_off) + _put_varint(d_size)
    i_off, i_size = emit(_build_block([(_short_successor(entries[-1][0]), handle)], restart_interval=1))
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


def _mlp_tensors(w: Weights) -> Dict[str, np.ndarray]:
    """{variable key: array} under the reference's naming, including the step aliasing (module docstring)."""
    g, p = "layer_with_weights-0", "layer_with_weights-1"
    out: Dict[str, np.ndarray] = {}

    def put(base: str, mlp: MLP):
        for i, (W, b) in enumerate(zip(mlp.W, mlp.b)):
            out[f"{base}/layer_set/{i}/kernel{_SUFFIX}"] = W
            out[f"{base}/layer_set/{i}/bias{_SUFFIX}"] = b

    for t in range(w.T - 1):
        put(f"{g}/message_fns/{t}", w.msg[t])
        put(f"{p}/pass_fns/{t}", w.pas[t])
    put(f"{g}/message_fn", w.msg[-1])
    put(f"{p}/pass_fn", w.pas[-1])
    put(f"{g}/update_fn", w.upd)
    return out


def save_weights(w: Weights, prefix: str, template_prefix: str):
    """Write ``w`` as a TF tensor-bundle checkpoint ``prefix`` (replaces ``model.save_weights``, reference
    ``charge_gn.py:462``), using the checkpoint ``template_prefix`` of the SAME architecture as the layout template:
    the Keras object graph (``_CHECKPOINTABLE_OBJECT_GRAPH``, which TensorFlow needs to restore by object), every key,
    shape, shard and offset are taken from the template; only the float tensors' bytes and CRCs are replaced.
    Re-saving an unmodified checkpoint reproduces the reference's files byte for byte (tests/test_checkpoint_io.py)."""
    index = read_index(template_prefix + ".index")
    if "" not in index:
        raise CheckpointError("template: bundle header entry missing")
    num_shards = 1
    for fnum, _, val in _proto_fields(index[""]):
        if fnum == 1:
            num_shards = val
    shards = {}
    for sid in range(num_shards):
        with open(f"{template_prefix}.data-{sid:05d}-of-{num_shards:05d}", "rb") as f:
            shards[sid] = bytearray(f.read())
    tensors = _mlp_tensors(w)
    entries: List[Tuple[bytes, bytes]] = []
    seen = set()
    for key in sorted(index, key=lambda k: k.encode("utf-8")):
        raw = index[key]
        ent = _parse_entry(raw) if key else None
        if ent is not None and ent.dtype == 1:
            if key not in tensors:
                raise CheckpointError(f"template variable {key} has no counterpart in the model (different architecture?)")
            arr = np.ascontiguousarray(tensors[key], dtype="<f4")
            if tuple(arr.shape) != ent.shape:
                raise CheckpointError(f"{key}: shape {arr.shape} does not match the template's {ent.shape}")
            blob = arr.tobytes()
            shards[ent.shard_id][ent.offset:ent.offset + ent.size] = blob
            # BundleEntryProto ends with field 6 (crc32c, fixed32: tag 0x35 + 4 bytes); everything before it is unchanged
            if len(raw) < 5 or raw[-5] != 0x35:
                raise CheckpointError(f"{key}: unexpected BundleEntryProto layout in the template")
            raw = raw[:-4] + struct.pack("<I", masked_crc32c(blob))
            seen.add(key)
        entries.append((key.encode("utf-8"), raw))
    missing = set(tensors) - seen
    if missing:
        raise CheckpointError(f"model variables absent from the template: {sorted(missing)[:3]} ...")
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    for sid, blob in shards.items():
        with open(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", "wb") as f:
            f.write(bytes(blob))
    _write_index(prefix + ".index", entries)
