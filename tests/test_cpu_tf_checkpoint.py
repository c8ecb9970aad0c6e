"""TensorFlow checkpoint-V2 reader (kfnet_b200/tf_checkpoint.py) vs a minimal writer that follows the same published
format (LevelDB table + BundleHeaderProto/BundleEntryProto).  No TensorFlow exists in the image, so this is a
round-trip / format-conformance test, not a test against real TF output."""
import os
import struct

import numpy as np
import pytest

from kfnet_b200 import tf_checkpoint as ck


# ---------------------------------------------------------------- minimal writer (test-only)
def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_varint(field, v): return _vi(field << 3) + _vi(v)
def _pb_bytes(field, b): return _vi((field << 3) | 2) + _vi(len(b)) + b
def _pb_fixed32(field, v): return _vi((field << 3) | 5) + struct.pack("<I", v)


def _block(entries, restart_interval=4):
    buf, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                shared += 1
        buf += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        last = k
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def write_bundle(prefix, tensors, block_entries=3):
    names = sorted(tensors)
    data, kv = bytearray(), [(b"", _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1)))]
    for n in names:
        a = np.asarray(tensors[n])                       # (ascontiguousarray would turn a scalar into shape (1,))
        raw = a.astype("<f4").tobytes()
        shape = b"".join(_pb_bytes(2, _pb_varint(1, d)) for d in a.shape)
        entry = (_pb_varint(1, ck.DT_FLOAT) + _pb_bytes(2, shape) + _pb_varint(3, 0) + _pb_varint(4, len(data)) +
                 _pb_varint(5, len(raw)) + _pb_fixed32(6, ck.mask_crc(ck.crc32c(raw))))
        kv.append((n.encode(), entry))
        data += raw
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    out, index_entries = bytearray(), []

    def emit(block):
        off = len(out)
        out.extend(block + b"\x00")
        out.extend(struct.pack("<I", ck.mask_crc(ck.crc32c(block + b"\x00"))))
        return _vi(off) + _vi(len(block))

    for i in range(0, len(kv), block_entries):
        chunk = kv[i:i + block_entries]
        index_entries.append((chunk[-1][0] + b"\xff", emit(_block(chunk))))     # separator key >= last key of the block
    meta = emit(_block([]))
    idx = emit(_block(index_entries, restart_interval=1))
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", ck.TABLE_MAGIC))
    open(prefix + ".index", "wb").write(bytes(out))


# ---------------------------------------------------------------- tests
def test_crc32c_known_answers():
    assert ck.crc32c(b"123456789") == 0xE3069283            # standard CRC-32C check value
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA            # RFC 3720 B.4 test vector
    assert ck.unmask_crc(ck.mask_crc(0x12345678)) == 0x12345678


def test_round_trip_and_lookup_by_name(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"ScoreNet/conv1a/kernel": rng.normal(size=(3, 3, 3, 64)).astype(np.float32),
               "ScoreNet/conv1a/bias": rng.normal(size=(64,)).astype(np.float32),
               "ScoreNet/conv1b/kernel": rng.normal(size=(3, 3, 64, 64)).astype(np.float32),
               "Temporal/fc1/kernel": rng.normal(size=(128, 64)).astype(np.float32),
               "Temporal/upconv2/kernel": rng.normal(size=(3, 3, 64, 128)).astype(np.float32),
               "global_step": np.array(7.0, np.float32)}
    prefix = str(tmp_path / "model.ckpt-7")
    write_bundle(prefix, tensors)
    header, entries = ck.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and set(entries) == set(tensors)
    assert entries["Temporal/upconv2/kernel"]["shape"] == (3, 3, 64, 128)
    got = ck.read_checkpoint(prefix, verify_data=True)
    for k, v in tensors.items():
        assert got[k].dtype == np.float32 and np.array_equal(got[k], v)
    only = ck.read_checkpoint(prefix, names=["Temporal/fc1/kernel"])
    assert list(only) == ["Temporal/fc1/kernel"]
    with pytest.raises(KeyError):
        ck.read_checkpoint(prefix, names=["ScoreNet/nope/kernel"])


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "model.ckpt-1")
    write_bundle(prefix, {"a/kernel": np.ones((4, 4), np.float32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[5] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(IOError):
        ck.read_index(prefix + ".index")
    open(prefix + ".index", "wb").write(b"not a table")
    with pytest.raises(IOError):
        ck.read_index(prefix + ".index")


def test_eval_snapshot_loader_accepts_tf_checkpoint(tmp_path, weights):
    from kfnet_b200 import eval as ev
    prefix = str(tmp_path / "model.ckpt-40000")
    write_bundle(prefix, weights, block_entries=16)
    snap, step = ev.get_snapshot(str(tmp_path))
    assert step == 40000 and snap == prefix
    w = ev.load_snapshot_weights(snap)
    assert set(w) == set(weights) and all(np.array_equal(w[k], weights[k]) for k in weights)
