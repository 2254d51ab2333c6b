"""TensorBoard event files without TensorFlow - the logging side of the training loop.

The reference builds three scalar summaries in its graph and logs them through two `tf.summary.FileWriter`s:
    self.train_writer = tf.summary.FileWriter(os.path.join(summaries_dir, 'train'))      src/linear_model.py:81
    self.test_writer  = tf.summary.FileWriter(os.path.join(summaries_dir, 'test'))       src/linear_model.py:82
    tf.summary.scalar('loss/loss', ...), ('loss/error_mm', ...), ('learning_rate/learning_rate', ...)   :130,134,148
    model.train_writer.add_summary(loss_summary, current_step)                           src/predict_3dpose.py:252-253
    model.test_writer.add_summary(summaries, current_step)                               src/predict_3dpose.py:296,323

Here `step()` returns `Summary(tag, value)` objects (a 2-tuple, so they unpack like before) that serialize to the
`Summary` protobuf a TF1 scalar summary is (`value { tag, simple_value }`), and `FileWriter` appends `Event` records
to `events.out.tfevents.<time>.<host>` in the TFRecord framing
    uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)
(CRC-32C from libp3d.so, the LevelDB masking of checkpoint.py).  The first record is the `file_version` event
"brain.Event:2".  Formats: tensorflow/core/util/event.proto, tensorflow/core/framework/summary.proto,
tensorflow/core/lib/io/record_writer.  Checked against the `tensorboard` package's own reader where it is installed
(tests/test_summary_cpu.py) - that reader, not TensorFlow, is the consumer these files are for.
"""
from __future__ import annotations

import collections
import os
import socket
import struct
import time

from .checkpoint import _pb_bytes, _pb_fields, _put_varint, crc32c, mask_crc

_SummaryBase = collections.namedtuple("Summary", ["tag", "value"])


class Summary(_SummaryBase):
    """One scalar summary: what `session.run(model.loss_summary)` hands to `add_summary` in the reference."""
    __slots__ = ()

    def SerializeToString(self):
        # Summary { repeated Value value = 1; }   Value { string tag = 1; float simple_value = 2; }
        v = _pb_bytes(1, self.tag.encode()) + _put_varint((2 << 3) | 5) + struct.pack("<f", float(self.value))
        return _pb_bytes(1, v)


def scalar(tag, value):
    """tf.summary.scalar(tag, tensor) evaluated: a Summary holding `value`."""
    return Summary(tag, value)


def _record(data):
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data)))


def _event(wall_time, step=None, file_version=None, summary=None):
    # Event { double wall_time = 1; int64 step = 2; oneof what { string file_version = 3; Summary summary = 5; } }
    out = _put_varint((1 << 3) | 1) + struct.pack("<d", wall_time)
    if step is not None:
        out += _put_varint(2 << 3) + _put_varint(int(step))
    if file_version is not None:
        out += _pb_bytes(3, file_version.encode())
    if summary is not None:
        out += _pb_bytes(5, summary)
    return out


class FileWriter(object):
    """tf.summary.FileWriter(logdir): creates the directory and the event file at once, appends one event per
    add_summary.  `summary` may be a Summary, its serialized bytes, or a list of Summary objects (the several values
    of a merged summary go into one event)."""
    _serial = 0

    def __init__(self, logdir, graph=None, flush_secs=120, filename_suffix=""):
        self._logdir = logdir
        self._flush_secs, self._last_flush = flush_secs, time.time()
        os.makedirs(logdir, exist_ok=True)
        now = time.time()
        FileWriter._serial += 1
        name = "events.out.tfevents.%010d.%s.%d.%d%s" % (int(now), socket.gethostname(), os.getpid(), FileWriter._serial,
                                                        filename_suffix)
        self.path = os.path.join(logdir, name)
        self._f = open(self.path, "ab")
        self._f.write(_record(_event(now, file_version="brain.Event:2")))
        self._f.flush()

    def get_logdir(self):
        return self._logdir

    def add_summary(self, summary, global_step=None):
        if self._f is None:
            raise RuntimeError("FileWriter is closed")
        if isinstance(summary, (bytes, bytearray)):
            blob = bytes(summary)
        elif isinstance(summary, Summary):
            blob = summary.SerializeToString()
        else:
            blob = b"".join(s.SerializeToString() for s in summary)
        self._f.write(_record(_event(time.time(), step=global_step, summary=blob)))
        if time.time() - self._last_flush >= self._flush_secs:
            self.flush()

    def add_graph(self, graph, global_step=None):
        """predict_3dpose.py:219 logs the TF graph; there is no graph here - accepted and ignored."""

    def flush(self):
        if self._f is not None:
            self._f.flush()
            self._last_flush = time.time()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_events(path, verify=True):
    """Parse an event file back: a list of dicts {wall_time, step, file_version | scalars: {tag: value}}.
    Raises ValueError on a CRC mismatch or a truncated record."""
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        if pos + 12 > len(buf):
            raise ValueError("truncated record header at byte %d" % pos)
        (n,), (hcrc,) = struct.unpack_from("<Q", buf, pos), struct.unpack_from("<I", buf, pos + 8)
        if verify and mask_crc(crc32c(buf[pos:pos + 8])) != hcrc:
            raise ValueError("length CRC mismatch at byte %d" % pos)
        if pos + 12 + n + 4 > len(buf):
            raise ValueError("truncated record at byte %d" % pos)
        data = buf[pos + 12:pos + 12 + n]
        (dcrc,) = struct.unpack_from("<I", buf, pos + 12 + n)
        if verify and mask_crc(crc32c(data)) != dcrc:
            raise ValueError("data CRC mismatch at byte %d" % pos)
        pos += 16 + n
        ev = {"wall_time": None, "step": 0}
        for fld, wt, v in _pb_fields(data):
            if fld == 1 and wt == 1:
                ev["wall_time"] = struct.unpack("<d", struct.pack("<Q", v))[0]
            elif fld == 2 and wt == 0:
                ev["step"] = v
            elif fld == 3 and wt == 2:
                ev["file_version"] = v.decode()
            elif fld == 5 and wt == 2:
                sc = ev.setdefault("scalars", collections.OrderedDict())
                for f1, w1, val in _pb_fields(v):
                    if f1 != 1 or w1 != 2:
                        continue
                    tag, x = None, None
                    for f2, w2, vv in _pb_fields(val):
                        if f2 == 1 and w2 == 2:
                            tag = vv.decode()
                        elif f2 == 2 and w2 == 5:
                            x = struct.unpack("<f", struct.pack("<I", vv))[0]
                    sc[tag] = x
        out.append(ev)
    return out
