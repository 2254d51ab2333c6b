"""TensorBoard event files behind model.train_writer / model.test_writer (src/linear_model.py:80-82,
:130-148; src/predict_3dpose.py:252-253,296,323).  CPU only: the writer is host code, CRC-32C comes from libp3d.so.

Pinned three ways: (1) bytes of one event assembled by hand from event.proto / summary.proto / the TFRecord framing,
(2) round trips and corruption detection through p3d.summary.read_events, (3) the `tensorboard` package's own reader
(the consumer these files are written for) when it is importable."""
import os
import struct

import numpy as np
import pytest

from p3d import summary
from p3d.checkpoint import crc32c, mask_crc


def test_scalar_summary_is_the_tf1_protobuf():
    s = summary.scalar("loss/loss", 0.5)
    tag, value = s                                        # still a 2-tuple (callers of step() unpack it)
    assert (tag, value) == ("loss/loss", 0.5)
    # Summary{ value{ tag: "loss/loss" simple_value: 0.5 } }: field 1 (len 16) { field 1 (len 9) "loss/loss", field 2 fixed32 }
    want = bytes([0x0A, 16, 0x0A, 9]) + b"loss/loss" + bytes([0x15]) + struct.pack("<f", 0.5)
    assert s.SerializeToString() == want


def test_event_file_bytes_by_hand(tmp_path):
    w = summary.FileWriter(str(tmp_path / "train"))
    w.add_summary(summary.scalar("a", 2.0), 7)
    w.close()
    raw = open(w.path, "rb").read()
    # record 0: Event{wall_time, file_version="brain.Event:2"}
    n0 = struct.unpack_from("<Q", raw, 0)[0]
    ev0 = raw[12:12 + n0]
    assert ev0[0] == 0x09 and ev0[9:] == bytes([0x1A, 13]) + b"brain.Event:2"
    assert struct.unpack_from("<I", raw, 8)[0] == mask_crc(crc32c(raw[:8]))
    assert struct.unpack_from("<I", raw, 12 + n0)[0] == mask_crc(crc32c(ev0))
    # record 1: Event{wall_time, step=7, summary{value{tag "a", 2.0}}}
    off = 16 + n0
    n1 = struct.unpack_from("<Q", raw, off)[0]
    ev1 = raw[off + 12:off + 12 + n1]
    sm = bytes([0x0A, 8, 0x0A, 1]) + b"a" + bytes([0x15]) + struct.pack("<f", 2.0)
    assert ev1[0] == 0x09 and ev1[9:] == bytes([0x10, 7, 0x2A, len(sm)]) + sm
    assert off + 16 + n1 == len(raw)
    assert os.path.basename(w.path).startswith("events.out.tfevents.")


def test_roundtrip_merged_bytes_and_corruption(tmp_path):
    w = summary.FileWriter(str(tmp_path))
    vals = np.random.RandomState(0).uniform(0, 3, 50).astype(np.float32)
    for i, v in enumerate(vals):
        w.add_summary(summary.scalar("loss/loss", v), i * 100)
    w.add_summary([summary.scalar("loss/loss", 1.5), summary.scalar("learning_rate/learning_rate", 1e-3)], 5000)
    w.add_summary(summary.scalar("loss/error_mm", 45.5).SerializeToString(), 2 ** 40)      # serialized bytes, 64-bit step
    w.add_graph(None)
    w.flush()
    w.close()
    with pytest.raises(RuntimeError):
        w.add_summary(summary.scalar("x", 0.0), 1)
    ev = summary.read_events(w.path)
    assert ev[0]["file_version"] == "brain.Event:2" and len(ev) == 53
    assert [e["step"] for e in ev[1:51]] == [i * 100 for i in range(50)]
    assert np.array_equal(np.array([e["scalars"]["loss/loss"] for e in ev[1:51]], dtype=np.float32), vals)
    assert list(ev[51]["scalars"].items()) == [("loss/loss", 1.5), ("learning_rate/learning_rate", np.float32(1e-3))]
    assert ev[52]["step"] == 2 ** 40 and ev[52]["scalars"] == {"loss/error_mm": 45.5}
    assert all(a["wall_time"] <= b["wall_time"] for a, b in zip(ev, ev[1:]))
    raw = bytearray(open(w.path, "rb").read())
    raw[40] ^= 1
    bad = str(tmp_path / "bad")
    open(bad, "wb").write(raw)
    with pytest.raises(ValueError):
        summary.read_events(bad)
    open(bad, "wb").write(bytes(raw[:-3]))
    with pytest.raises(ValueError):
        summary.read_events(bad)


def test_tensorboard_reads_the_files(tmp_path):
    loader_mod = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    w = summary.FileWriter(str(tmp_path / "log" / "train"))
    want = []
    for step in range(1, 21):
        loss, lr = 1.0 / step, 1e-3 * 0.96 ** (step / 100000.0)
        w.add_summary(summary.scalar("loss/loss", loss), step)
        w.add_summary(summary.scalar("learning_rate/learning_rate", lr), step)
        want += [(step, "loss/loss", np.float32(loss)), (step, "learning_rate/learning_rate", np.float32(lr))]
    w.close()
    got, versions = [], []
    for e in loader_mod.LegacyEventFileLoader(w.path).Load():
        if e.HasField("file_version"):
            versions.append(e.file_version)
        for v in e.summary.value:
            got.append((e.step, v.tag, np.float32(v.simple_value)))
    assert versions == ["brain.Event:2"]
    assert got == want
    # and through the accumulator TensorBoard's scalar dashboard is fed from
    acc_mod = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
    acc = acc_mod.EventAccumulator(str(tmp_path / "log" / "train"))
    acc.Reload()
    assert sorted(acc.Tags()["scalars"]) == ["learning_rate/learning_rate", "loss/loss"]
    sc = acc.Scalars("loss/loss")
    assert [s.step for s in sc] == list(range(1, 21))
    assert np.allclose([s.value for s in sc], [1.0 / s for s in range(1, 21)], rtol=1e-6)
