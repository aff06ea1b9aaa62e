"""Scalar summaries for TensorBoard without TensorFlow - the step AFTER the path.

The reference logs four scalars per update (actorcritic/examples/atari/a2c_acktr.py:83-96,112-114,128-133):

    with tf.name_scope('model'):
        tf.summary.scalar('policy_loss', objective.policy_loss)          # -> tag "model/policy_loss"
        tf.summary.scalar('baseline_loss', objective.baseline_loss)
        tf.summary.scalar('policy_entropy', objective.mean_entropy)
    with tf.name_scope('environment'):
        tf.summary.scalar('episode_reward', episode_reward_placeholder)
    summary_op = tf.summary.merge_all()
    summary_writer = tf.summary.FileWriter(summary_path, session.graph)
    summary, step, _ = session.run([summary_op, global_step, optimize_op], feed_dict={...})
    summary_writer.add_summary(summary, step)

This module keeps those names.  `scalar()` records (tag, source) pairs, `merge_all()` returns a fetch token that
`Session.run` evaluates together with the train step (the loss scalars come from the same 64-byte read-back as every other
fetch; placeholders such as the episode reward are taken from the feed_dict), and `FileWriter` writes standard
`events.out.tfevents.*` files: TFRecord framing (length, masked CRC-32C of the length, payload, masked CRC-32C of the
payload) around hand-encoded `Event` protocol buffers, so TensorBoard reads them as it reads TensorFlow's.
"""
import contextlib
import os
import socket
import struct
import threading
import time

import numpy as np

from .session import Fetch, Placeholder

# ------------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli), the checksum of the TFRecord format
# ------------------------------------------------------------------------------------------------
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data):
    crc = 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord(payload):
    header = struct.pack("<Q", len(payload))
    return header + struct.pack("<I", masked_crc32c(header)) + payload + struct.pack("<I", masked_crc32c(payload))


# ------------------------------------------------------------------------------------------------
# protocol buffers by hand (tensorflow/core/util/event.proto, framework/summary.proto)
#   Event   { double wall_time = 1; int64 step = 2; oneof what { string file_version = 3; Summary summary = 5; } }
#   Summary { repeated Value value = 1; }     Value { string tag = 1; float simple_value = 2; }
# ------------------------------------------------------------------------------------------------
def _varint(n):
    n &= (1 << 64) - 1          # int64 fields: two's complement, 10 bytes when negative
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_delimited(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_summary(values):
    """values: iterable of (tag, float) -> serialized Summary."""
    out = b""
    for tag, value in values:
        v = _len_delimited(1, tag.encode("utf-8")) + _varint((2 << 3) | 5) + struct.pack("<f", float(value))
        out += _len_delimited(1, v)
    return out


def encode_event(wall_time, step=None, summary=None, file_version=None):
    out = _varint((1 << 3) | 1) + struct.pack("<d", float(wall_time))
    if step is not None:
        out += _varint((2 << 3) | 0) + _varint(int(step))
    if file_version is not None:
        out += _len_delimited(3, file_version.encode("utf-8"))
    if summary is not None:
        out += _len_delimited(5, summary)
    return out


# ------------------------------------------------------------------------------------------------
# the tf.summary surface the reference uses
# ------------------------------------------------------------------------------------------------
_scope = threading.local()
_collection = []          # (tag, Fetch | Placeholder) in registration order: tf.GraphKeys.SUMMARIES of the default graph


def _prefix():
    return "/".join(getattr(_scope, "names", []))


@contextlib.contextmanager
def name_scope(name):
    """tf.name_scope for summary tags."""
    names = getattr(_scope, "names", [])
    _scope.names = names + [name]
    try:
        yield
    finally:
        _scope.names = names


def placeholder(dtype=np.float32, shape=(), name="summary_value"):
    """tf.placeholder for values that only exist on the host (a2c_acktr.py:80: the mean episode reward)."""
    return Placeholder(name, dtype, list(shape))


def scalar(name, tensor):
    """tf.summary.scalar: `tensor` is a fetch token of this framework (objective.policy_loss ...) or a placeholder."""
    if not isinstance(tensor, (Fetch, Placeholder)):
        raise TypeError("summary.scalar needs a fetch token or a placeholder, got %r" % (tensor,))
    tag = (_prefix() + "/" + name) if _prefix() else name
    _collection.append((tag, tensor))
    return SummaryFetch([(tag, tensor)])


def reset_default_collection():
    """Forget the registered summaries (tf.reset_default_graph)."""
    del _collection[:]


class SummaryFetch(Fetch):
    """What tf.summary.merge_all() returns: `Session.run` evaluates the sources and yields a serialized Summary."""

    def __init__(self, items):
        owner = next((t.owner for _, t in items if isinstance(t, Fetch)), None)
        super().__init__("summary", owner, "summary")
        self.items = list(items)

    def sources(self):
        return [t for _, t in self.items if isinstance(t, Fetch)]

    def build(self, values, feed):
        """values: {id(fetch): scalar}; feed: the feed_dict (placeholders)."""
        pairs = []
        for tag, src in self.items:
            if isinstance(src, Placeholder):
                if src not in feed:
                    raise ValueError("placeholder %r of summary %r must be fed" % (src.name, tag))
                pairs.append((tag, float(np.asarray(feed[src]).reshape(()))))
            else:
                pairs.append((tag, float(values[id(src)])))
        return encode_summary(pairs)


def merge_all():
    """tf.summary.merge_all(): every summary registered so far (None when there is none, like TensorFlow)."""
    return SummaryFetch(_collection) if _collection else None


def no_op():
    """tf.no_op(): what the reference fetches instead of the summary op when summaries are off (a2c_acktr.py:94-96);
    `Session.run` returns None for it and `FileWriter.add_summary(None, step)` ignores it."""
    return Fetch("no_op", None, "no_op")


def merge(summaries):
    """tf.summary.merge."""
    items = []
    for s in summaries:
        items.extend(s.items)
    return SummaryFetch(items)


class FileWriter:
    """tf.summary.FileWriter(logdir, graph=None): `add_summary(summary, global_step)`, `flush()`, `close()`.  The graph
    argument is accepted and ignored (there is no graph)."""

    def __init__(self, logdir, graph=None, filename_suffix=""):
        os.makedirs(logdir, exist_ok=True)
        self._path = os.path.join(logdir, "events.out.tfevents.%010d.%s%s" % (int(time.time()), socket.gethostname(),
                                                                              filename_suffix))
        self._file = open(self._path, "ab")
        self._file.write(tfrecord(encode_event(time.time(), file_version="brain.Event:2")))
        self._file.flush()

    path = property(lambda self: self._path)

    def get_logdir(self):
        return os.path.dirname(self._path)

    def add_summary(self, summary, global_step=None):
        if summary is None:                      # tf.no_op() stand-in when summaries are off (a2c_acktr.py:94-96)
            return
        if not isinstance(summary, (bytes, bytearray)):
            raise TypeError("add_summary expects a serialized Summary (what Session.run returns for a summary op)")
        self._file.write(tfrecord(encode_event(time.time(), step=global_step, summary=bytes(summary))))

    def flush(self):
        self._file.flush()

    def close(self):
        if not self._file.closed:
            self._file.flush()
            self._file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def read_events(path):
    """Parse an event file back (tests, offline inspection): yields dicts {wall_time, step, file_version, scalars}.
    Verifies both CRCs of every record."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        header = data[pos:pos + 8]
        (length,) = struct.unpack("<Q", header)
        (hcrc,) = struct.unpack("<I", data[pos + 8:pos + 12])
        payload = data[pos + 12:pos + 12 + length]
        (pcrc,) = struct.unpack("<I", data[pos + 12 + length:pos + 16 + length])
        if hcrc != masked_crc32c(header) or pcrc != masked_crc32c(payload) or len(payload) != length:
            raise ValueError("corrupt record at byte %d of %s" % (pos, path))
        pos += 16 + length
        yield _decode_event(payload)


def _read_varint(buf, pos):
    shift = value = 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _read_varint(buf, pos)
        elif wire == 1:
            value, pos = buf[pos:pos + 8], pos + 8
        elif wire == 5:
            value, pos = buf[pos:pos + 4], pos + 4
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            value, pos = buf[pos:pos + n], pos + n
        else:
            raise ValueError("unsupported wire type %d" % wire)
        yield field, wire, value


def _decode_event(payload):
    ev = {"wall_time": None, "step": 0, "file_version": None, "scalars": {}}
    for field, wire, value in _fields(payload):
        if field == 1 and wire == 1:
            ev["wall_time"] = struct.unpack("<d", value)[0]
        elif field == 2 and wire == 0:
            ev["step"] = value - (1 << 64) if value >= (1 << 63) else value
        elif field == 3 and wire == 2:
            ev["file_version"] = bytes(value).decode("utf-8")
        elif field == 5 and wire == 2:
            for f1, w1, v1 in _fields(value):
                if f1 == 1 and w1 == 2:
                    tag, simple = None, None
                    for f2, w2, v2 in _fields(v1):
                        if f2 == 1 and w2 == 2:
                            tag = bytes(v2).decode("utf-8")
                        elif f2 == 2 and w2 == 5:
                            simple = struct.unpack("<f", v2)[0]
                    ev["scalars"][tag] = simple
    return ev
