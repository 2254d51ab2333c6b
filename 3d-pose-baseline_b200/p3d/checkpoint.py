"""TensorFlow checkpoint (V2 "tensor bundle") reader / writer without TensorFlow - the data format on the weight side
of the hot path: what `tf.train.Saver(tf.global_variables())` of the reference writes and restores
(src/linear_model.py:151; src/predict_3dpose.py:165-181 restore, :328 save).

A checkpoint `<prefix>` is
    <prefix>.index                 a LevelDB-format sorted string table: key "" -> BundleHeaderProto,
                                   key <variable name> -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
    <prefix>.data-00000-of-00001   the raw little-endian tensor bytes, back to back in key order
plus the text file `checkpoint` in the directory naming the latest prefix (CheckpointState).

TensorFlow is not part of the reference tree (third-party, un-vendored, version unpinned - README.md:23) and is not
installable here, and the reference ships no checkpoint: this module follows the published formats (LevelDB table
format; tensorflow/core/protobuf/tensor_bundle.proto; tensorflow/core/util/tensor_bundle) and is tested by round trips,
CRC known answers and hand-assembled tables (tests/test_checkpoint_cpu.py) - PARITY UNPINNED against real TF files.

Variable names are the TF1 graph's (see include/p3d.h): "linear_model/w1", ".../batch_normalization/gamma",
"linear_model/two_linear_0/w2_0", ..., Adam slots "<var>/Adam" and "<var>/Adam_1", "beta1_power", "beta2_power",
"global_step" (int32 scalar), "learning_rate" (the un-decayed base rate, linear_model.py:86).
"""
from __future__ import annotations

import os
import struct

import numpy as np

from ._lib import lib

TABLE_MAGIC = 0xDB4775248B80FB57
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 19: np.dtype("<f2")}
_DT_OF = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9, np.dtype("float16"): 19}


# ------------------------------------------------------------------ crc32c (Castagnoli), LevelDB masking
def crc32c(data, init=0):
    """CRC-32C of a bytes-like object / contiguous NumPy array (native slice-by-8 in libp3d.so, host code)."""
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(lib.p3d_crc32c(a.ctypes.data, a.nbytes, init))
    b = bytes(data)
    return int(lib.p3d_crc32c(b, len(b), init))


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / minimal protobuf
def _put_varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _pb_fields(buf):
    """Yield (field number, wire type, value) of one serialized message (varint, 64-bit, bytes, 32-bit)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _encode_entry(dtype_enum, shape, offset, size, crc_masked):
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto.dim=2 {size=1}), shard_id=3, offset=4, size=5, crc32c=6 (fixed32)."""
    shp = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)
    out = _pb_varint(1, dtype_enum) + _pb_bytes(2, shp)
    if offset:
        out += _pb_varint(4, offset)
    out += _pb_varint(5, size) + _put_varint((6 << 3) | 5) + struct.pack("<I", crc_masked)
    return out


def _decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------ snappy (only if a table block is compressed)
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


# ------------------------------------------------------------------ LevelDB table format
def _block_entries(block):
    """(key, value) pairs of one table block (prefix-compressed keys; the restart array is not needed to scan)."""
    if len(block) < 4:
        raise ValueError("table block too short")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    if end < 0:
        raise ValueError("corrupt table block")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared]); pos += non_shared
        yield key, bytes(block[pos:pos + vlen]); pos += vlen


def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype = raw[:size], raw[size]
    if verify:
        want = struct.unpack_from("<I", raw, size + 1)[0]
        if mask_crc(crc32c(raw[:size + 1])) != want:
            raise ValueError("table block checksum mismatch")
    if ctype == 0:
        return body
    if ctype == 1:
        return _snappy_decompress(body)
    raise ValueError("unknown table block compression %d" % ctype)


def read_table(path, verify=True):
    """All (key bytes, value bytes) of a LevelDB-format table file, in key order."""
    out = []
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        n = f.tell()
        if n < 48:
            raise ValueError("%s is too short to be a table" % path)
        f.seek(n - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
            raise ValueError("%s: bad table magic (not a TensorFlow checkpoint index)" % path)
        pos = 0
        _, pos = _get_varint(footer, pos); _, pos = _get_varint(footer, pos)       # metaindex handle
        ioff, pos = _get_varint(footer, pos); isz, pos = _get_varint(footer, pos)   # index handle
        for _, handle in _block_entries(_read_block(f, ioff, isz, verify)):
            boff, p2 = _get_varint(handle, 0)
            bsz, _ = _get_varint(handle, p2)
            out.extend(_block_entries(_read_block(f, boff, bsz, verify)))
    return out


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_table(path, items, block_size=262144):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order.  Uncompressed blocks (what
    tensorflow's BundleWriter asks for), restart interval 16, masked CRC-32C trailers."""
    with open(path, "wb") as f:
        offset = 0

        def emit(block):
            nonlocal offset
            trailer = b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))
            f.write(block + trailer)
            handle = _put_varint(offset) + _put_varint(len(block))
            offset += len(block) + 5
            return handle

        index, bb, prev = [], _BlockBuilder(), None
        for key, value in items:
            if prev is not None and key <= prev:
                raise ValueError("table keys must be strictly increasing")
            bb.add(key, value)
            prev = key
            if len(bb.buf) >= block_size:
                index.append((bb.last, emit(bb.finish())))
                bb = _BlockBuilder()
        if bb.count or not index:
            index.append((bb.last, emit(bb.finish())))
        meta_handle = emit(_BlockBuilder().finish())
        ib = _BlockBuilder(restart_interval=1)
        for key, handle in index:
            ib.add(key, handle)
        index_handle = emit(ib.finish())
        footer = meta_handle + index_handle
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))


# ------------------------------------------------------------------ tensor bundle
def _data_path(prefix, shard=0, num_shards=1):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def list_variables(prefix):
    """[(name, shape, numpy dtype)] of a checkpoint, like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + ".index"):
        if k:
            e = _decode_entry(v)
            out.append((k.decode(), tuple(e["shape"]), _DT.get(e["dtype"])))
    return out


def read_bundle(prefix, verify=True):
    """{variable name: ndarray} of the checkpoint `<prefix>` (.index + .data-*).  Checks the per-tensor CRC-32C."""
    if not os.path.isfile(prefix + ".index"):
        raise ValueError("Asked to load checkpoint {0}, but it does not seem to exist".format(prefix))
    entries = read_table(prefix + ".index", verify)
    num_shards = 1
    for k, v in entries:
        if k == b"":
            for f, _, val in _pb_fields(v):
                if f == 1:
                    num_shards = val
                elif f == 2 and val != 0:
                    raise ValueError("big-endian checkpoints are not supported")
    files, out = {}, {}
    try:
        for k, v in entries:
            if k == b"":
                continue
            e = _decode_entry(v)
            if e["sliced"]:
                raise ValueError("partitioned variable %s is not supported" % k.decode())
            if e["dtype"] not in _DT:
                raise ValueError("variable %s has unsupported dtype enum %d" % (k.decode(), e["dtype"]))
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open(_data_path(prefix, sid, num_shards), "rb")
            files[sid].seek(e["offset"])
            raw = files[sid].read(e["size"])
            dt = _DT[e["dtype"]]
            count = int(np.prod(e["shape"])) if e["shape"] else 1
            if len(raw) != e["size"] or e["size"] != count * dt.itemsize:
                raise ValueError("variable %s: size mismatch in the data file" % k.decode())
            a = np.frombuffer(raw, dtype=dt).reshape(e["shape"])
            if verify and e["crc32c"] is not None and mask_crc(crc32c(a)) != e["crc32c"]:
                raise ValueError("variable %s: checksum mismatch" % k.decode())
            out[k.decode()] = a
    finally:
        for f in files.values():
            f.close()
    return out


def write_bundle(prefix, tensors):
    """Write {name: ndarray} as one-shard checkpoint `<prefix>` (float32/float64/int32/int64/float16 tensors)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors.keys(), key=lambda s: s.encode())
    # BundleHeaderProto: num_shards=1 (field 1), endianness LITTLE=0 (default, omitted), version {producer=1} (field 3)
    items = [(b"", _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))]
    offset = 0
    with open(_data_path(prefix), "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])
            if a.dtype not in _DT_OF:
                raise ValueError("variable %s: dtype %s cannot be stored" % (n, a.dtype))
            shape = a.shape                                   # ascontiguousarray would turn a scalar into [1]
            a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False))
            f.write(a.tobytes())
            items.append((n.encode(), _encode_entry(_DT_OF[np.dtype(a.dtype.name)], shape, offset, a.nbytes, mask_crc(crc32c(a)))))
            offset += a.nbytes
    write_table(prefix + ".index", items)
    return prefix


# ------------------------------------------------------------------ the `checkpoint` state file
def get_checkpoint_state(train_dir, latest_filename="checkpoint"):
    """tf.train.get_checkpoint_state (predict_3dpose.py:165): None, or a dict with `model_checkpoint_path` and
    `all_model_checkpoint_paths` (relative paths are resolved against train_dir, as TF does)."""
    path = os.path.join(train_dir, latest_filename)
    if not os.path.isfile(path):
        return None
    state = {"model_checkpoint_path": None, "all_model_checkpoint_paths": []}
    with open(path) as f:
        for line in f:
            if ":" not in line:
                continue
            k, v = line.split(":", 1)
            v = v.strip().strip('"')
            if not os.path.isabs(v):
                v = os.path.join(train_dir, v)
            if k.strip() == "model_checkpoint_path":
                state["model_checkpoint_path"] = v
            elif k.strip() == "all_model_checkpoint_paths":
                state["all_model_checkpoint_paths"].append(v)
    return state if state["model_checkpoint_path"] else None


def update_checkpoint_state(train_dir, prefix, max_to_keep=10, latest_filename="checkpoint"):
    """What Saver.save does after writing: record `prefix` as the latest checkpoint, keep the newest `max_to_keep`
    (linear_model.py:151) and delete the files of older ones."""
    st = get_checkpoint_state(train_dir, latest_filename)
    paths = [p for p in (st["all_model_checkpoint_paths"] if st else []) if os.path.abspath(p) != os.path.abspath(prefix)]
    paths.append(prefix)
    while max_to_keep and len(paths) > max_to_keep:
        old = paths.pop(0)
        for suffix in (".index", ".data-00000-of-00001", ".meta"):
            if os.path.isfile(old + suffix):
                os.remove(old + suffix)
    rel = [os.path.relpath(p, train_dir) for p in paths]
    with open(os.path.join(train_dir, latest_filename), "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % rel[-1])
        for r in rel:
            f.write('all_model_checkpoint_paths: "%s"\n' % r)
