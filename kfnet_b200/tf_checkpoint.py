"""Reader for TensorFlow checkpoint-V2 ("tensor bundle") files: `model.ckpt-<step>.index` + `.data-00000-of-0000N`.

Replaces, for inference, what `tf.train.Saver.restore` does in the reference (KFNet/train.py:317-321, eval.py:66-68):
variables are looked up BY NAME (`ScoreNet/conv1a/kernel`, `Temporal/fc1/bias`, ...) and returned as
float32 numpy arrays in TF layout, ready for `Engine(weights)`.

Format (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, which is LevelDB's table format):
  * `.index` is an immutable sorted table: data blocks, a metaindex block, an index block and a 48-byte footer
    (two BlockHandles = varint64 offset + varint64 size, zero padding to 40 bytes, magic 0xdb4775248b80fb57 LE).
    Every block is followed by a 5-byte trailer (compression type, masked crc32c).  Block = prefix-compressed entries
    (varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value) + uint32 restart offsets + uint32 count.
  * key ""  -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
    key name -> BundleEntryProto  {1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                                   6: crc32c (fixed32), 7: slices}
  * tensor bytes live in `<prefix>.data-<shard:05d>-of-<num_shards:05d>` at [offset, offset+size), little-endian.

STATUS: written from the format definition above; no TensorFlow (and no real checkpoint) exists in the build image, so
it is verified only by the round-trip test in tests/test_cpu_tf_checkpoint.py, whose writer follows the same
definition.  Snappy-compressed index blocks (not what BundleWriter emits) and sliced (partitioned) variables raise.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}

# ---------------------------------------------------------------------------------------------- crc32c (Castagnoli)
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """leveldb/TF crc masking: rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf
def _varint(buf: bytes, pos: int):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _proto_fields(buf: bytes):
    """Yields (field_number, wire_type, value) of one protobuf message; value is int or bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _parse_shape(buf: bytes):
    dims = []
    for f, _, v in _proto_fields(buf):
        if f == 2:                                  # repeated Dim
            size = 0
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf: bytes):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, _, v in _proto_fields(buf):
        if f == 1: e["dtype"] = v
        elif f == 2: e["shape"] = _parse_shape(v)
        elif f == 3: e["shard_id"] = v
        elif f == 4: e["offset"] = v
        elif f == 5: e["size"] = v
        elif f == 6: e["crc32c"] = v
        elif f == 7: e["sliced"] = True
    return e


# ---------------------------------------------------------------------------------------------- table reading
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = unmask_crc(struct.unpack_from("<I", data, offset + size + 1)[0])
        if crc32c(data[offset:offset + size + 1]) != stored:
            raise IOError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise NotImplementedError("checkpoint index block is compressed (type %d); BundleWriter writes uncompressed tables" % ctype)
    return block


def _block_entries(block: bytes):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str, verify: bool = True):
    """-> (header dict, {name: entry dict}) of a `.index` file."""
    data = open(index_path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise IOError("%s is not a TensorFlow checkpoint-V2 index (bad table magic)" % index_path)
    footer = data[-48:]
    _, p = _varint(footer, 0); _, p = _varint(footer, p)            # metaindex handle (unused)
    ioff, p = _varint(footer, p); isz, p = _varint(footer, p)       # index handle
    entries, header = {}, None
    for _, handle in _block_entries(_read_block(data, ioff, isz, verify)):
        boff, q = _varint(handle, 0); bsz, q = _varint(handle, q)
        for key, value in _block_entries(_read_block(data, boff, bsz, verify)):
            if key == b"":
                header = {f: v for f, _, v in _proto_fields(value) if f in (1, 2)}
            else:
                entries[key.decode()] = _parse_entry(value)
    if header is None:
        raise IOError("checkpoint index has no bundle header")
    return {"num_shards": header.get(1, 1), "endianness": header.get(2, 0)}, entries


def read_checkpoint(prefix: str, names=None, verify: bool = True, verify_data: bool = False) -> dict:
    """Reads variables of a checkpoint prefix (`.../model.ckpt-<step>`) -> {name: float32/int ndarray in TF layout}.
    `names`: iterable of variable names to load (default: all); missing names raise KeyError.
    `verify` checks the index blocks' checksums; `verify_data` also checks every tensor's crc32c (pure-Python, slow)."""
    header, entries = read_index(prefix + ".index", verify)
    if header["endianness"] != 0:
        raise NotImplementedError("big-endian checkpoints are not supported")
    wanted = list(entries) if names is None else list(names)
    files, out = {}, {}
    for name in wanted:
        e = entries[name]
        if e["sliced"]:
            raise NotImplementedError("partitioned variable %s" % name)
        if e["dtype"] not in _DTYPES:
            if names is None:
                continue                                              # e.g. string tensors of the Saver itself
            raise NotImplementedError("dtype %d of %s" % (e["dtype"], name))
        shard = e["shard_id"]
        if shard not in files:
            files[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, header["num_shards"]), "rb")
        f = files[shard]
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise IOError("short read for %s" % name)
        if verify_data and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise IOError("checksum mismatch for %s" % name)
        arr = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"])
        out[name] = arr.astype(np.float32) if e["dtype"] == DT_FLOAT else arr.copy()
    for f in files.values():
        f.close()
    return out


def load_kfnet_weights(prefix: str) -> dict:
    """The variables eval restores: scopes 'ScoreNet' and 'Temporal' (KFNet/eval.py:66-68), by the names of SURVEY App. C."""
    from .synth import weight_shapes
    shapes = weight_shapes()
    w = read_checkpoint(prefix, names=list(shapes))
    for k, shp in shapes.items():
        if tuple(w[k].shape) != tuple(shp):
            raise ValueError("variable %s has shape %s in the checkpoint, expected %s" % (k, w[k].shape, shp))
    return w
