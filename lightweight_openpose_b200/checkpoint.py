"""TensorBundle (TF checkpoint V2) reader for ``model/model.ckpt-61236`` and a
seeded synthetic data-shard generator with the same layout.

The reference restores its variables with ``tf.train.Saver().restore(sess, prefix)``
(reference ``test&eval/model_test.py:43,49``, ``test&eval/model_json.py:53,60``).
Without TensorFlow the same bytes are read here directly:

* ``<prefix>.index`` is a LevelDB-style sorted table (uncompressed blocks with
  prefix-compressed keys, a 48-byte footer ending in the magic
  ``0xdb4775248b80fb57``).  Key ``""`` holds a ``BundleHeaderProto``; every
  other key is a variable name whose value is a ``BundleEntryProto``
  ``{1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size, 6: crc32c}``.
* ``<prefix>.data-00000-of-00001`` is the concatenation of the raw
  little-endian tensors at those offsets.

The shipped repository only carries the ``.index`` file (the data shard is a
stripped large blob), so :func:`synthetic_variables` produces a deterministic
stand-in with exactly the shapes the index lists; the loader path downstream is
identical for real and synthetic weights.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT = 1
DT_INT64 = 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT64: np.dtype("<i8"), 3: np.dtype("<i4")}

# Number of bytes the full (training) checkpoint's data shard must have
# (last entry ``lr`` at offset 48 706 256, 4 bytes).
EXPECTED_SHARD_BYTES = 48_706_260


@dataclass(frozen=True)
class BundleEntry:
    name: str
    dtype: int
    shape: Tuple[int, ...]
    shard_id: int
    offset: int
    size: int
    crc32c: Optional[int]  # masked, as stored

    @property
    def np_dtype(self) -> np.dtype:
        return _DTYPES[self.dtype]


# ----------------------------------------------------------------------------
# varint / protobuf helpers
# ----------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _proto_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for a serialized protobuf."""
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims: List[int] = []
    for field, _, val in _proto_fields(buf):
        if field == 2:  # repeated Dim dim
            size = 0
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(name: str, buf: bytes) -> BundleEntry:
    dtype = 0
    shape: Tuple[int, ...] = ()
    shard = offset = size = 0
    crc = None
    for field, _, val in _proto_fields(buf):
        if field == 1:
            dtype = val
        elif field == 2:
            shape = _parse_shape(val)
        elif field == 3:
            shard = val
        elif field == 4:
            offset = val
        elif field == 5:
            size = val
        elif field == 6:
            crc = val
    return BundleEntry(name, dtype, shape, shard, offset, size, crc)


# ----------------------------------------------------------------------------
# table reader
# ----------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int) -> List[Tuple[bytes, bytes]]:
    """Decode one table block into (key, value) pairs.  The block is followed
    by a 1-byte compression type (must be 0) and a 4-byte CRC."""
    if data[offset + size] != 0:
        raise ValueError("compressed index blocks are not supported")
    block = data[offset:offset + size]
    num_restarts = struct.unpack_from("<I", block, size - 4)[0]
    limit = size - 4 - 4 * num_restarts
    out: List[Tuple[bytes, bytes]] = []
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(index_path: str) -> Tuple[Dict[str, BundleEntry], dict]:
    """Parse ``<prefix>.index``.  Returns ``(entries, header)`` where ``entries``
    maps variable name -> :class:`BundleEntry` in table (sorted-key) order."""
    with open(index_path, "rb") as fh:
        data = fh.read()
    if len(data) < 48:
        raise ValueError("index file too small")
    footer = data[-48:]
    magic = struct.unpack_from("<Q", footer, 40)[0]
    if magic != TABLE_MAGIC:
        raise ValueError(f"bad table magic {magic:#x}")
    pos = 0
    _meta_off, pos = _varint(footer, pos)
    _meta_size, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)

    entries: Dict[str, BundleEntry] = {}
    header: dict = {}
    for _, handle in _read_block(data, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _read_block(data, boff, bsize):
            if key == b"":
                for field, _, val in _proto_fields(value):
                    if field == 1:
                        header["num_shards"] = val
                    elif field == 2:
                        header["endianness"] = val
                continue
            name = key.decode("utf-8")
            entries[name] = _parse_entry(name, value)
    header.setdefault("num_shards", 1)
    return entries, header


def is_inference_variable(name: str) -> bool:
    """True for the 173 tensors the forward graph reads (optimizer slots and
    training scalars are skipped)."""
    if name.endswith("/Adam") or name.endswith("/Adam_1"):
        return False
    return name.split("/")[0] in ("BackBone", "Cpm", "InitialStage", "RefinementStage")


def inference_entries(entries: Dict[str, BundleEntry]) -> Dict[str, BundleEntry]:
    return {k: v for k, v in entries.items() if is_inference_variable(k)}


# ----------------------------------------------------------------------------
# CRC32C (Castagnoli), TF masking
# ----------------------------------------------------------------------------
_CRC_TABLE: Optional[np.ndarray] = None


def _crc_table() -> np.ndarray:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tbl = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tbl[i] = c
        _CRC_TABLE = tbl
    return _CRC_TABLE


_CRC_SLICES: Optional[np.ndarray] = None


def _crc_slices() -> np.ndarray:
    """Slice-by-8 tables: T[k][b] = CRC state after byte b followed by k zero bytes."""
    global _CRC_SLICES
    if _CRC_SLICES is None:
        t0 = _crc_table().astype(np.uint64)
        t = np.zeros((8, 256), dtype=np.uint64)
        t[0] = t0
        for k in range(1, 8):
            t[k] = (t[k - 1] >> np.uint64(8)) ^ t0[(t[k - 1] & np.uint64(0xFF)).astype(np.int64)]
        _CRC_SLICES = t
    return _CRC_SLICES


def _crc32c_native(data: bytes) -> Optional[int]:
    """lwop_crc32c from the in-tree library when it is built (hardware-speed slice-by-8 in C)."""
    try:
        from . import _lib
        import ctypes as C
        lib = _lib.load()
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        return int(lib.lwop_crc32c(buf, len(data)))
    except Exception:
        return None


def crc32c(data: bytes, native: bool = True) -> int:
    """CRC32C of ``data``.  Uses the C implementation in ``_lwop.so`` when available (a 48 MB shard in well under a
    second); otherwise a numpy slice-by-8 over 8-byte words (the per-byte Python loop took 0.6 s per MiB)."""
    if native and len(data) >= 1024:
        v = _crc32c_native(data)
        if v is not None:
            return v
    tbl = _crc_table()
    crc = 0xFFFFFFFF
    n8 = len(data) // 8
    if n8 >= 64:
        t = _crc_slices()
        words = np.frombuffer(data, dtype="<u8", count=n8)
        # the CRC recurrence is sequential in the state, so the word loop stays in Python, but each step is 8 table
        # look-ups on scalars instead of 8 interpreted byte iterations
        tl = [t[k].tolist() for k in range(8)]
        for w in words.tolist():
            x = w ^ crc
            crc = (tl[7][x & 0xFF] ^ tl[6][(x >> 8) & 0xFF] ^ tl[5][(x >> 16) & 0xFF] ^ tl[4][(x >> 24) & 0xFF] ^
                   tl[3][(x >> 32) & 0xFF] ^ tl[2][(x >> 40) & 0xFF] ^ tl[1][(x >> 48) & 0xFF] ^ tl[0][(x >> 56) & 0xFF])
        data = data[n8 * 8:]
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """TensorFlow's masked CRC: rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------
# data shard access
# ----------------------------------------------------------------------------
def data_shard_path(prefix: str, shard: int = 0, num_shards: int = 1) -> str:
    return f"{prefix}.data-{shard:05d}-of-{num_shards:05d}"


def load_variables(prefix: str, verify_crc: bool = False,
                   only_inference: bool = True) -> Dict[str, np.ndarray]:
    """Read variables from ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``.
    Raises FileNotFoundError if the data shard is missing (it is in the shipped
    reference: ``.MISSING_LARGE_BLOBS``)."""
    entries, header = read_index(prefix + ".index")
    if header.get("endianness", 0) not in (0, None):
        raise ValueError("big-endian TensorBundle shards are not supported")
    path = data_shard_path(prefix, 0, header["num_shards"])
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as fh:
        for name, e in entries.items():
            if only_inference and not is_inference_variable(name):
                continue
            if e.shard_id != 0:
                raise ValueError(f"{name}: lives in shard {e.shard_id}; only single-shard bundles are supported")
            fh.seek(e.offset)
            raw = fh.read(e.size)
            if len(raw) != e.size:
                raise ValueError(f"{name}: data shard truncated")
            if verify_crc and e.crc32c is not None:
                if mask_crc(crc32c(raw)) != e.crc32c:
                    raise ValueError(f"{name}: crc32c mismatch")
            out[name] = np.frombuffer(raw, dtype=e.np_dtype).reshape(e.shape).copy()
    return out


def apply_calibration(raw: Dict[str, np.ndarray], gains: Dict[str, np.ndarray],
                      shifts: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Per output channel ``c`` of each layer: pre-activation -> ``g_c * pre + s_c``
    (through gamma/beta for BN layers, kernel/bias otherwise)."""
    from .graph import LAYERS, KIND_SEPARABLE
    out = dict(raw)
    for layer in LAYERS:
        g = gains.get(layer.name)
        if g is None:
            continue
        g = np.asarray(g, dtype=np.float64)
        s = np.asarray(shifts[layer.name], dtype=np.float64)
        if layer.bn:
            out[layer.bn + "/gamma"] = (out[layer.bn + "/gamma"] * g).astype(np.float32)
            out[layer.bn + "/beta"] = (out[layer.bn + "/beta"] * g + s).astype(np.float32)
        else:
            k = layer.name + ("/pointwise_kernel" if layer.kind == KIND_SEPARABLE else "/kernel")
            out[k] = (out[k].astype(np.float64) * g).astype(np.float32)   # last axis = out channel
            if layer.bias:
                b = layer.name + "/bias"
                out[b] = (out[b].astype(np.float64) * g + s).astype(np.float32)
    return out


SYNTH_RANDOM_MIX = 0.5   # weight of the He-random channel mixing next to the identity path

_CALIBRATION_CACHE: Dict[int, Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]] = {}


def _load_calibration(seed: int):
    if seed not in _CALIBRATION_CACHE:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                            f"synth_calibration_{seed}.npz")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path}: no calibration for seed {seed}; run tools/calibrate_synthetic.py "
                "or pass calibrated=False")
        z = np.load(path)
        gains = {k[5:]: z[k] for k in z.files if k.startswith("gain:")}
        shifts = {k[6:]: z[k] for k in z.files if k.startswith("shift:")}
        _CALIBRATION_CACHE[seed] = (gains, shifts)
    return _CALIBRATION_CACHE[seed]


def synthetic_variables(entries: Dict[str, BundleEntry], seed: int = 61236,
                        calibrated: bool = True) -> Dict[str, np.ndarray]:
    """Deterministic stand-in weights with the index's shapes.  Conv kernels are
    He-normal ``N(0, 2/fan_in)`` across channels times a jittered binomial 3x3
    spatial stencil, biases ``N(0, 0.01)``, BN ``gamma~U(0.5,1.5)``,
    ``beta~N(0,0.1)``, ``moving_mean~N(0,0.1)``, ``moving_variance~U(0.5,1.5)``
    (SURVEY.md 8(d) C1).  Variables are drawn in sorted-name order so the result
    does not depend on dict order.  With ``calibrated`` the per-channel gains and
    shifts in ``data/synth_calibration_<seed>.npz`` (computed once by
    ``tools/calibrate_synthetic.py``) make every hidden pre-activation zero-mean /
    unit-std on synthetic frames, as in a trained BN network, and put the heads
    at heat ~ N(-0.12, 0.1^2), paf ~ N(0, 0.15^2).  The raw recipe grows
    activations to ~1e4 and collapses onto its DC path, against which the
    absolute parity tolerances would be meaningless."""
    if calibrated:
        raw = synthetic_variables(entries, seed, calibrated=False)
        gains, shifts = _load_calibration(seed)
        return apply_calibration(raw, gains, shifts)
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name in sorted(entries):
        if not is_inference_variable(name):
            continue
        e = entries[name]
        leaf = name.rsplit("/", 1)[1]
        shape = e.shape
        if leaf in ("kernel", "pointwise_kernel", "depthwise_kernel"):
            # channel mixing is He-random; the 3x3 spatial profile is a jittered
            # low-pass (binomial) stencil, as trained filters mostly are -- iid
            # spatial taps act as random high-pass filters that destroy the
            # image signal and leave only amplified rounding noise.
            kh, kw, cin, cout = shape
            fan_in = cin if leaf != "depthwise_kernel" else 1
            mix = SYNTH_RANDOM_MIX * rng.normal(0.0, np.sqrt(2.0 / fan_in), size=(1, 1, cin, cout))
            if leaf != "depthwise_kernel" and cin >= 32:
                # identity-like backbone path (channel i feeds channels i mod cin):
                # trained networks are contractive to perturbations, a purely
                # random critical ReLU stack amplifies rounding noise ~1.05x/layer.
                eye = np.zeros((cin, cout))
                if cout >= cin:
                    eye[np.arange(cout) % cin, np.arange(cout)] = 1.0
                else:
                    eye[np.arange(cin), np.arange(cin) % cout] = np.sqrt(cout / cin)
                mix = mix + eye.reshape(1, 1, cin, cout)
            if kh == 3:
                prof = np.outer([1.0, 2.0, 1.0], [1.0, 2.0, 1.0]) / 16.0
                jitter = 1.0 + 0.3 * rng.normal(0.0, 1.0, size=shape)
                v = prof.reshape(3, 3, 1, 1) * jitter * mix
            else:
                v = np.broadcast_to(mix, shape).copy()
        elif leaf == "bias":
            v = rng.normal(0.0, 0.01, size=shape)
        elif leaf == "gamma":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf in ("beta", "moving_mean"):
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf == "moving_variance":
            v = rng.uniform(0.5, 1.5, size=shape)
        else:
            raise ValueError(f"unexpected variable {name}")
        out[name] = v.astype(np.float32)
    return out


def write_synthetic_shard(prefix_index: str, out_path: str, seed: int = 61236) -> int:
    """Write a data shard laid out exactly per the index (offsets/sizes) filled
    with :func:`synthetic_variables`; non-inference entries are zero.  Returns
    the number of bytes written."""
    entries, _ = read_index(prefix_index + ".index")
    total = max(e.offset + e.size for e in entries.values())
    buf = bytearray(total)
    syn = synthetic_variables(entries, seed)
    for name, arr in syn.items():
        e = entries[name]
        raw = arr.astype(e.np_dtype).tobytes()
        assert len(raw) == e.size, name
        buf[e.offset:e.offset + e.size] = raw
    if "global_step" in entries:
        e = entries["global_step"]
        buf[e.offset:e.offset + e.size] = np.array(61236, dtype="<i8").tobytes()
    with open(out_path, "wb") as fh:
        fh.write(buf)
    return total


def default_index_prefix() -> str:
    """The checkpoint prefix shipped with this repo (index only)."""
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), "model", "model.ckpt-61236")


def load_or_synthesize(prefix: Optional[str] = None, seed: int = 61236,
                       verify_crc: bool = True) -> Tuple[Dict[str, np.ndarray], str]:
    """Real weights if the data shard exists, else the seeded synthetic set.
    Returns ``(variables, source)`` with source in {"checkpoint", "synthetic"}."""
    prefix = prefix or default_index_prefix()
    entries, header = read_index(prefix + ".index")
    if os.path.exists(data_shard_path(prefix, 0, header["num_shards"])):
        return load_variables(prefix, verify_crc=verify_crc), "checkpoint"
    return synthetic_variables(entries, seed), "synthetic"
