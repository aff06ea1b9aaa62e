"""Checkpoint / resume on the GPU: save after a few updates, restore into a fresh model in a fresh session, and the
continued run is bit-identical to the uninterrupted one (parameters, optimizer slots, K-FAC sums, inverses, counters)."""
import numpy as np
import pytest
import torch

import synth
from test_gpu_api import _build

pytestmark = pytest.mark.gpu


def _feed(model, batch):
    return {model.observations_placeholder: batch["observations"],
            model.bootstrap_observations_placeholder: batch["bootstrap_observations"],
            model.actions_placeholder: batch["actions"], model.rewards_placeholder: batch["rewards"],
            model.terminals_placeholder: batch["terminals"]}


@pytest.mark.parametrize("acktr", [True, False])
def test_resume_is_bit_identical(acktr, tmp_path):
    from actorcritic_b200 import checkpoint
    e_count, t_count, total, cut = 4, 5, 7, 4          # cold 2 -> covariances -> inverses every update (see _build)
    batches = [synth.rollout(500 + u, e_count, t_count, 4, obs_kind="sparse") for u in range(total)]
    fisher = [synth.fisher_samples(600 + u, e_count * t_count) for u in range(total)]

    def inject(session, u):
        y_hat, eps = fisher[u]
        session.fisher_injection = (torch.from_numpy(y_hat).cuda(), torch.from_numpy(eps).cuda()) if acktr else None

    # uninterrupted run, checkpoint after `cut` updates
    ac, model, objective, global_step, optimize_op = _build(acktr, e_count, t_count, seed=5)
    saver = checkpoint.Saver()                          # tf.train.Saver(): the most recently built model
    with ac.Session() as session:
        for u in range(total):
            inject(session, u)
            step, _ = session.run([global_step, optimize_op], feed_dict=_feed(model, batches[u]))
            if u == cut - 1:
                path = saver.save(session, str(tmp_path / "atari"), step)
                assert path.endswith("atari-%d.npz" % step)
        want = checkpoint.state_to_arrays(model.engine)
    assert checkpoint.latest_checkpoint(str(tmp_path)) == path

    # the file holds the reference's variable names and layouts
    with np.load(path) as z:
        assert z["conv1/weights"].shape == (8, 8, 4, 32) and z["fc_policy/bias"].shape == (4,)
        assert ("conv2/weights/velocity" in z.files) == acktr and ("conv2/weights/RMSProp" in z.files) == (not acktr)
        assert ("kfac/cov/A/heads" in z.files) == acktr

    # fresh model + session: restore BEFORE the first train step (the reference's order, a2c_acktr.py:100-104)
    ac2, model2, objective2, global_step2, optimize_op2 = _build(acktr, e_count, t_count, seed=99)
    saver2 = checkpoint.Saver(model2)
    with ac2.Session() as session:
        saver2.restore(session, checkpoint.latest_checkpoint(str(tmp_path)))
        for u in range(cut, total):
            inject(session, u)
            step, _ = session.run([global_step2, optimize_op2], feed_dict=_feed(model2, batches[u]))
        got = checkpoint.state_to_arrays(model2.engine)
    assert int(got["global_step"]) == int(want["global_step"]) == step
    for k in want:
        assert np.array_equal(got[k], want[k]), k

    # restore into a live learner as well
    with ac2.Session() as session:
        saver2.restore(session, path)
        assert global_step2.eval() == int(np.load(path)["global_step"])
        with pytest.raises(FileNotFoundError):
            saver2.restore(session, str(tmp_path / "nope.npz"))


def test_restore_rejects_a_different_architecture(tmp_path):
    from actorcritic_b200 import checkpoint
    ac, model, objective, global_step, optimize_op = _build(True, 2, 2)
    batch = synth.rollout(1, 2, 2, 4, obs_kind="sparse")
    with ac.Session() as session:
        session.run(optimize_op, feed_dict=_feed(model, batch))
        path = checkpoint.Saver(model).save(session, str(tmp_path / "m"), 1)
    ac2, model2, objective2, gs2, op2 = _build(False, 2, 2)      # conv3 = 64
    with ac2.Session() as session:
        session.run(op2, feed_dict=_feed(model2, batch))
        with pytest.raises(ValueError):
            checkpoint.Saver(model2).restore(session, path)


def test_restore_then_act_then_save_keeps_the_optimizer_state(tmp_path):
    """restore -> sample_actions (builds an acting-only engine) -> save -> restore -> train: the K-FAC running sums,
    inverses, velocities and counters of the first checkpoint survive (the reference's Ctrl-C handler can save during the
    first rollout of a resumed run, a2c_acktr.py:135-143), and `session.run(global_step)` right after the restore
    reports the restored step."""
    from actorcritic_b200 import checkpoint
    e_count, t_count = 4, 5
    batches = [synth.rollout(700 + u, e_count, t_count, 4, obs_kind="sparse") for u in range(6)]
    y_hat, eps = synth.fisher_samples(800, e_count * t_count)
    fisher = (torch.from_numpy(y_hat).cuda(), torch.from_numpy(eps).cuda())
    ac, model, objective, global_step, optimize_op = _build(True, e_count, t_count, seed=5)
    with ac.Session() as session:
        session.fisher_injection = fisher
        for u in range(4):
            step, _ = session.run([global_step, optimize_op], feed_dict=_feed(model, batches[u]))
        first = checkpoint.Saver(model).save(session, str(tmp_path / "a"), step)
        want_state = checkpoint.state_to_arrays(model.engine)
        session.run(optimize_op, feed_dict=_feed(model, batches[4]))
        want_after = checkpoint.state_to_arrays(model.engine)

    ac2, model2, objective2, global_step2, optimize_op2 = _build(True, e_count, t_count, seed=99)
    saver2 = checkpoint.Saver(model2)
    with ac2.Session() as session:
        session.fisher_injection = fisher
        saver2.restore(session, first)
        assert int(session.run(global_step2)) == int(want_state["global_step"])      # before any engine exists
        obs = batches[0]["observations"][:, :1]
        model2.sample_actions(obs, session)                                           # acting-only engine
        assert model2.engine is not None and not model2.engine.is_learner
        second = saver2.save(session, str(tmp_path / "b"), int(want_state["global_step"]))
        with np.load(second) as z:
            assert bool(z["meta/acktr"]) and "kfac/cov/A/heads" in z.files and "conv2/weights/velocity" in z.files
            for k in want_state:
                assert np.array_equal(z[k], want_state[k]), k
    ac3, model3, objective3, global_step3, optimize_op3 = _build(True, e_count, t_count, seed=7)
    with ac3.Session() as session:
        session.fisher_injection = fisher
        checkpoint.Saver(model3).restore(session, second)
        model3.sample_actions(batches[0]["observations"][:, :1], session)            # acting engine first, learner after
        session.run(optimize_op3, feed_dict=_feed(model3, batches[4]))
        got = checkpoint.state_to_arrays(model3.engine)
    for k in want_after:
        assert np.array_equal(got[k], want_after[k]), k
