"""TensorFlow checkpoint (tensor bundle) reader/writer of p3d/checkpoint.py - SURVEY 8(f).2.  CPU only (host code).
No TF checkpoint exists in the reference tree and TensorFlow is not installable: PARITY UNPINNED against real TF files;
what is checked here are the published building blocks (CRC-32C known answers, LevelDB masking, varints, block/prefix
coding, footer magic) plus round trips."""
import os
import struct

import numpy as np
import pytest

from p3d import checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 (iSCSI) appendix B.4 test patterns and the usual check value
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # chaining, unaligned starts, NumPy input
    data = np.random.RandomState(0).randint(0, 256, 10007).astype(np.uint8)
    whole = ck.crc32c(data)
    for cut in (0, 1, 7, 8, 9, 5000, 10007):
        assert ck.crc32c(data[cut:].tobytes(), ck.crc32c(data[:cut].tobytes())) == whole
    assert ck.crc32c(data[3:]) == ck.crc32c(data[3:].tobytes())


def test_leveldb_crc_mask():
    # leveldb crc32c_test: Mask is not the identity, not an involution, and Unmask inverts it
    crc = ck.crc32c(b"foo")
    assert ck.mask_crc(crc) != crc and ck.mask_crc(ck.mask_crc(crc)) != crc
    assert ck.unmask_crc(ck.mask_crc(crc)) == crc
    assert ck.unmask_crc(ck.unmask_crc(ck.mask_crc(ck.mask_crc(crc)))) == crc
    assert ck.mask_crc(0) == 0xA282EAD8


def test_varint_roundtrip():
    for n in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 35 + 5, 2 ** 63 + 1):
        b = ck._put_varint(n)
        assert ck._get_varint(b, 0) == (n, len(b))
    assert ck._put_varint(300) == b"\xac\x02"


def test_table_layout_by_hand(tmp_path):
    """One tiny table assembled byte by byte from the LevelDB format description, read back by read_table."""
    def block(entries):
        body = b""
        for shared, key_delta, value in entries:
            body += bytes([shared, len(key_delta), len(value)]) + key_delta + value
        return body + struct.pack("<II", 0, 1)

    def with_trailer(b):
        return b + b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(b + b"\x00")))

    data = block([(0, b"apple", b"1"), (4, b"y", b"22"), (0, b"b", b"")])         # apple, apply (shared "appl"), b
    meta = block([])
    index = block([(0, b"b", bytes([0, len(data)]))])                              # handle: offset 0, size
    off_meta = len(data) + 5
    off_index = off_meta + len(meta) + 5
    footer = bytes([off_meta, len(meta), off_index, len(index)])
    blob = with_trailer(data) + with_trailer(meta) + with_trailer(index) + footer + bytes(40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    p = tmp_path / "t.index"
    p.write_bytes(blob)
    assert ck.read_table(str(p)) == [(b"apple", b"1"), (b"apply", b"22"), (b"b", b"")]
    # a flipped bit is caught by the block checksum
    bad = bytearray(blob); bad[4] ^= 1
    p.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        ck.read_table(str(p))
    # our writer produces the same bytes for the same content
    q = tmp_path / "w.index"
    ck.write_table(str(q), [(b"apple", b"1"), (b"apply", b"22"), (b"b", b"")])
    assert q.read_bytes() == blob


def test_table_many_blocks_and_order(tmp_path):
    items = [(("key%06d" % i).encode(), os.urandom(i % 50)) for i in range(5000)]
    p = str(tmp_path / "big.index")
    ck.write_table(p, items, block_size=4096)
    assert ck.read_table(p) == items
    with pytest.raises(ValueError):
        ck.write_table(p, [(b"b", b""), (b"a", b"")])


def test_snappy_block_decoder():
    # literal "abcd" + copy(offset 4, length 8) -> "abcdabcdabcd" ; the length prefix is a varint
    comp = bytes([12, (3 << 2) | 0]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert ck._snappy_decompress(comp) == b"abcdabcdabcd"


def test_entry_proto_bytes():
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:32} dim{size:1024}}, offset: 5, size: 131072, crc32c: 0x01020304}
    b = ck._encode_entry(1, (32, 1024), 5, 131072, 0x01020304)
    assert b == bytes([0x08, 0x01, 0x12, 0x09, 0x12, 0x02, 0x08, 0x20, 0x12, 0x03, 0x08, 0x80, 0x08,
                       0x20, 0x05, 0x28, 0x80, 0x80, 0x08, 0x35, 0x04, 0x03, 0x02, 0x01])
    e = ck._decode_entry(b)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"]) == (1, [32, 1024], 5, 131072, 0x01020304)


def test_bundle_roundtrip_and_corruption(tmp_path):
    rng = np.random.RandomState(1)
    t = {"linear_model/w1": rng.normal(size=(32, 1024)).astype(np.float32),
         "linear_model/b1": rng.normal(size=1024).astype(np.float32),
         "linear_model/w1/Adam": np.zeros((32, 1024), np.float32),
         "global_step": np.int32(4874200), "learning_rate": np.float32(1e-3),
         "half": rng.normal(size=(3, 5)).astype(np.float16), "i64": np.arange(7, dtype=np.int64)}
    prefix = str(tmp_path / "checkpoint-4874200")
    ck.write_bundle(prefix, t)
    assert os.path.isfile(prefix + ".index") and os.path.isfile(prefix + ".data-00000-of-00001")
    back = ck.read_bundle(prefix)
    assert sorted(back) == sorted(t)
    for k in t:
        assert back[k].dtype == np.asarray(t[k]).dtype and np.array_equal(back[k], t[k])
    names = [n for n, _, _ in ck.list_variables(prefix)]
    assert names == sorted(names, key=lambda s: s.encode())
    assert dict((n, s) for n, s, _ in ck.list_variables(prefix))["linear_model/w1"] == (32, 1024)
    # header entry: num_shards = 1, version.producer = 1
    assert ck.read_table(prefix + ".index")[0] == (b"", bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01]))
    # a flipped byte in the data file is caught by the tensor checksum
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read()); d[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(ValueError, match="checksum"):
        ck.read_bundle(prefix)
    with pytest.raises(ValueError, match="does not seem to exist"):
        ck.read_bundle(str(tmp_path / "checkpoint-1"))


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert ck.get_checkpoint_state(d) is None
    for step in range(1, 14):
        prefix = os.path.join(d, "checkpoint-%d" % step)
        ck.write_bundle(prefix, {"global_step": np.int32(step)})
        ck.update_checkpoint_state(d, prefix, max_to_keep=10)
    st = ck.get_checkpoint_state(d)
    assert st["model_checkpoint_path"] == os.path.join(d, "checkpoint-13")
    assert len(st["all_model_checkpoint_paths"]) == 10
    assert not os.path.exists(os.path.join(d, "checkpoint-3.index")) and os.path.exists(os.path.join(d, "checkpoint-4.index"))
    txt = open(os.path.join(d, "checkpoint")).read().splitlines()
    assert txt[0] == 'model_checkpoint_path: "checkpoint-13"' and txt[1] == 'all_model_checkpoint_paths: "checkpoint-4"'


def test_bundle_protos_against_the_protobuf_runtime_and_tf_descriptors():
    """The index entries as Google's protobuf runtime sees them.  TensorShapeProto, VersionDef and the DataType enum are
    TensorFlow's own generated descriptors (vendored by the `tensorboard` package); BundleHeaderProto / BundleEntryProto
    are declared here from tensorflow/core/protobuf/tensor_bundle.proto (tensorboard does not ship that one)."""
    pytest.importorskip("google.protobuf")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="p3d_test/tensor_bundle.proto", package="p3d_test", syntax="proto3")
    fd.dependency.extend([shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    hdr = fd.message_type.add(name="BundleHeaderProto")
    en = hdr.enum_type.add(name="Endianness")
    en.value.add(name="LITTLE", number=0); en.value.add(name="BIG", number=1)
    hdr.field.add(name="num_shards", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    hdr.field.add(name="endianness", number=2, type=F.TYPE_ENUM, label=F.LABEL_OPTIONAL, type_name=".p3d_test.BundleHeaderProto.Endianness")
    hdr.field.add(name="version", number=3, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name="." + versions_pb2.VersionDef.DESCRIPTOR.full_name)
    ent = fd.message_type.add(name="BundleEntryProto")
    ent.field.add(name="dtype", number=1, type=F.TYPE_ENUM, label=F.LABEL_OPTIONAL, type_name="." + types_pb2.DESCRIPTOR.enum_types_by_name["DataType"].full_name)
    ent.field.add(name="shape", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name="." + shape_pb2.TensorShapeProto.DESCRIPTOR.full_name)
    ent.field.add(name="shard_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    ent.field.add(name="offset", number=4, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    ent.field.add(name="size", number=5, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    ent.field.add(name="crc32c", number=6, type=F.TYPE_FIXED32, label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.Default()
    fdesc = pool.Add(fd) if hasattr(pool, "Add") else pool.AddSerializedFile(fd.SerializeToString())
    if fdesc is None:
        fdesc = pool.FindFileByName(fd.name)
    get = getattr(message_factory, "GetMessageClass", None)
    Header = get(fdesc.message_types_by_name["BundleHeaderProto"])
    Entry = get(fdesc.message_types_by_name["BundleEntryProto"])

    # dtype enum numbers of the reader/writer tables are TensorFlow's
    for np_dt, name in ((np.float32, "DT_FLOAT"), (np.float64, "DT_DOUBLE"), (np.int32, "DT_INT32"), (np.int64, "DT_INT64"),
                        (np.float16, "DT_HALF")):
        assert ck._DT_OF[np.dtype(np_dt)] == types_pb2.DataType.Value(name)
        assert ck._DT[types_pb2.DataType.Value(name)] == np.dtype(np_dt).newbyteorder("<")

    # our bytes -> protobuf runtime
    for shape, off in (((1024, 48), 0), ((32, 1024), 123456789012), ((), 8), ((7,), 0)):
        raw = ck._encode_entry(1, shape, off, 4 * int(np.prod(shape)) if shape else 4, 0xDEADBEEF)
        e = Entry.FromString(raw)
        assert e.dtype == types_pb2.DT_FLOAT and [d.size for d in e.shape.dim] == list(shape)
        assert (e.shard_id, e.offset, e.size, e.crc32c) == (0, off, 4 * int(np.prod(shape)) if shape else 4, 0xDEADBEEF)
        assert not e.shape.unknown_rank
    # protobuf runtime -> our decoder (the bytes TensorFlow itself would put into the index)
    e = Entry(dtype=types_pb2.DT_INT32, shard_id=0, offset=77, size=4, crc32c=0x01020304)
    d = ck._decode_entry(e.SerializeToString())
    assert (d["dtype"], d["shape"], d["offset"], d["size"], d["crc32c"]) == (3, [], 77, 4, 0x01020304)
    e = Entry(dtype=types_pb2.DT_HALF, shape=shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=1024),
                                                                            shape_pb2.TensorShapeProto.Dim(size=1024)]),
              offset=1 << 33, size=2 << 20, crc32c=0xFFFFFFFF)
    d = ck._decode_entry(e.SerializeToString())
    assert (d["dtype"], d["shape"], d["offset"], d["size"], d["crc32c"]) == (19, [1024, 1024], 1 << 33, 2 << 20, 0xFFFFFFFF)


def test_written_index_header_parses_as_bundle_header(tmp_path):
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    prefix = ck.write_bundle(os.path.join(str(tmp_path), "c"), {"a": np.zeros(3, np.float32)})
    items = dict(ck.read_table(prefix + ".index"))
    hdr = items[b""]
    # num_shards = 1 (field 1 varint), version (field 3) = VersionDef{producer: 1}
    fields = list(ck._pb_fields(hdr))
    assert fields[0] == (1, 0, 1)
    v = versions_pb2.VersionDef.FromString(fields[1][2])
    assert fields[1][0] == 3 and v.producer == 1 and v.min_consumer == 0 and list(v.bad_consumers) == []
