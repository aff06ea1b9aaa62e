"""Summary writer (a2c_acktr.py:83-96,112-114,128-133) without a GPU: CRC-32C known answers, the TFRecord
framing and the hand-encoded Event / Summary protocol buffers (checked with google.protobuf against descriptors built from
TensorFlow's published .proto definitions), the tf.summary-shaped surface, and the episode-reward aggregation."""
import struct

import numpy as np
import pytest

from actorcritic_b200 import summary
from actorcritic_b200.session import Fetch


def test_crc32c_known_answers():
    assert summary.crc32c(b"") == 0
    assert summary.crc32c(b"123456789") == 0xE3069283            # the standard check value of CRC-32C
    assert summary.crc32c(bytes(32)) == 0x8A9136AA                # RFC 3720 B.4
    assert summary.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert summary.crc32c(bytes(range(32))) == 0x46DD794E


def test_tfrecord_framing():
    rec = summary.tfrecord(b"payload")
    (length,) = struct.unpack("<Q", rec[:8])
    assert length == 7 and rec[12:19] == b"payload" and len(rec) == 8 + 4 + 7 + 4
    assert struct.unpack("<I", rec[8:12])[0] == summary.masked_crc32c(rec[:8])
    assert struct.unpack("<I", rec[19:])[0] == summary.masked_crc32c(b"payload")


def _event_message_classes():
    """Event / Summary message classes from descriptors written after tensorflow/core/util/event.proto and
    tensorflow/core/framework/summary.proto (the fields this framework emits)."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="acx_test_event.proto", package="acxtest", syntax="proto3")
    value = fd.message_type.add(name="Value")
    value.field.add(name="tag", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_STRING,
                    label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    value.field.add(name="simple_value", number=2, type=descriptor_pb2.FieldDescriptorProto.TYPE_FLOAT,
                    label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    summ = fd.message_type.add(name="Summary")
    summ.field.add(name="value", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_MESSAGE, type_name=".acxtest.Value",
                   label=descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED)
    ev = fd.message_type.add(name="Event")
    ev.field.add(name="wall_time", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_DOUBLE,
                 label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    ev.field.add(name="step", number=2, type=descriptor_pb2.FieldDescriptorProto.TYPE_INT64,
                 label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    ev.field.add(name="file_version", number=3, type=descriptor_pb2.FieldDescriptorProto.TYPE_STRING,
                 label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    ev.field.add(name="summary", number=5, type=descriptor_pb2.FieldDescriptorProto.TYPE_MESSAGE, type_name=".acxtest.Summary",
                 label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        factory = message_factory.MessageFactory(pool)
        return factory.GetPrototype(pool.FindMessageTypeByName("acxtest.Event"))
    return get(pool.FindMessageTypeByName("acxtest.Event"))


def test_event_encoding_parses_with_protobuf():
    Event = _event_message_classes()
    payload = summary.encode_event(1234.5, step=77, summary=summary.encode_summary([("model/policy_loss", -0.25),
                                                                                     ("environment/episode_reward", 31.0)]))
    ev = Event()
    ev.ParseFromString(payload)
    assert ev.wall_time == 1234.5 and ev.step == 77
    assert [(v.tag, v.simple_value) for v in ev.summary.value] == [("model/policy_loss", -0.25),
                                                                   ("environment/episode_reward", 31.0)]
    head = Event()
    head.ParseFromString(summary.encode_event(1.0, file_version="brain.Event:2"))
    assert head.file_version == "brain.Event:2"
    neg = Event()
    neg.ParseFromString(summary.encode_event(0.0, step=-3))
    assert neg.step == -3
    # and protobuf's own serialisation of the same message is byte-identical
    assert ev.SerializeToString() == payload


def test_tf_summary_surface_and_file_roundtrip(tmp_path):
    summary.reset_default_collection()
    assert summary.merge_all() is None
    losses = {name: Fetch(name, object()) for name in ("policy_loss", "baseline_loss", "mean_entropy")}
    episode_reward = summary.placeholder(np.float32, [])
    with summary.name_scope("model"):                              # a2c_acktr.py:84-88
        summary.scalar("policy_loss", losses["policy_loss"])
        summary.scalar("baseline_loss", losses["baseline_loss"])
        summary.scalar("policy_entropy", losses["mean_entropy"])
    with summary.name_scope("environment"):                        # :89-90
        summary.scalar("episode_reward", episode_reward)
    op = summary.merge_all()
    assert [t for t, _ in op.items] == ["model/policy_loss", "model/baseline_loss", "model/policy_entropy",
                                        "environment/episode_reward"]
    assert op.sources() == [losses["policy_loss"], losses["baseline_loss"], losses["mean_entropy"]]
    with pytest.raises(TypeError):
        summary.scalar("bad", 3.0)
    values = {id(losses["policy_loss"]): 0.5, id(losses["baseline_loss"]): 2.0, id(losses["mean_entropy"]): 1.25}
    with pytest.raises(ValueError):
        op.build(values, {})                                       # the placeholder must be fed
    with summary.FileWriter(str(tmp_path), graph=None) as writer:
        for step in (10, 20):
            writer.add_summary(op.build(values, {episode_reward: np.float32(step * 1.5)}), step)
        writer.add_summary(None, 30)                               # summaries switched off: tf.no_op() yields None
        writer.flush()
        path = writer.path
    events = list(summary.read_events(path))
    assert events[0]["file_version"] == "brain.Event:2" and len(events) == 3
    assert events[1]["step"] == 10 and events[2]["step"] == 20
    assert events[2]["scalars"] == {"model/policy_loss": 0.5, "model/baseline_loss": 2.0, "model/policy_entropy": 1.25,
                                    "environment/episode_reward": 30.0}
    with open(path, "r+b") as f:                                   # a flipped payload byte is caught by the CRC
        f.seek(-6, 2)
        f.write(b"\x00")
    with pytest.raises(ValueError):
        list(summary.read_events(path))
    summary.reset_default_collection()


def test_episode_info_wrapper_and_reward_aggregation():
    from actorcritic_b200.envs.atari.wrappers import EpisodeInfoWrapper

    class Env:
        def __init__(self):
            self.t = 0

        def reset(self):
            return 0

        def step(self, action):
            self.t += 1
            return self.t, float(self.t), self.t % 3 == 0, {"lives": 1}

    env = EpisodeInfoWrapper(Env())
    env.reset()
    infos = [env.step(0)[3] for _ in range(6)]
    assert "episode" not in infos[0] and infos[2]["episode"]["total_reward"] == 6.0 and infos[5]["episode"]["total_reward"] == 15.0
    assert infos[2]["lives"] == 1
    batch = [infos, [{} for _ in range(6)]]
    r = EpisodeInfoWrapper.get_episode_rewards_from_info_batch(batch)
    assert r.shape == (2, 6) and r.dtype == np.float32 and r[0, 2] == 6.0 and r[0, 5] == 15.0
    assert np.isnan(r[0, 0]) and np.all(np.isnan(r[1]))
    assert np.nanmean(r) == 10.5                                   # a2c_acktr.py:112-114
    assert np.all(np.isnan(EpisodeInfoWrapper.get_episode_rewards_from_info_batch([[{}]])))


def test_session_expands_summary_sources_and_no_op_without_a_device(monkeypatch):
    """Host logic of Session.run for summary ops: the summary's sources join the work list once, the serialized Summary is
    assembled from the same evaluation, no_op yields None (a2c_acktr.py:117-126 fetches [summary_op, global_step, op])."""
    from actorcritic_b200.session import Session
    summary.reset_default_collection()
    owner = object()
    policy_loss, entropy, step = Fetch("policy_loss", owner), Fetch("mean_entropy", owner), Fetch("global_step", None)
    reward = summary.placeholder(np.float32, [])
    with summary.name_scope("model"):
        summary.scalar("policy_loss", policy_loss)
        summary.scalar("policy_entropy", entropy)
    with summary.name_scope("environment"):
        summary.scalar("episode_reward", reward)
    op = summary.merge_all()
    seen = []

    def fake_plain(self, flist, feed):
        seen.append([f.kind for f in flist])
        return {id(f): {"policy_loss": 0.25, "mean_entropy": 1.5, "global_step": 42}[f.kind] for f in flist}

    monkeypatch.setattr(Session, "_evaluate_plain", fake_plain)
    session = Session.__new__(Session)                       # no CUDA device needed for the host logic
    out = session.run([op, step, policy_loss, summary.no_op()], feed_dict={reward: 3.0})
    assert seen == [["global_step", "policy_loss", "mean_entropy"]]          # policy_loss evaluated once, entropy added
    assert out[1] == 42 and out[2] == 0.25 and out[3] is None
    ev = summary._decode_event(summary.encode_event(0.0, step=out[1], summary=out[0]))
    assert ev["scalars"] == {"model/policy_loss": 0.25, "model/policy_entropy": 1.5, "environment/episode_reward": 3.0}
    assert session.run(summary.no_op()) is None and seen == [["global_step", "policy_loss", "mean_entropy"]]
    summary.reset_default_collection()
