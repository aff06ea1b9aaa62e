"""The drop-in boundary on the GPU: the reference's own call sequence (actorcritic/examples/atari/a2c_acktr.py:48-126,
218-253; docs/guide.rst:120-154) written against actorcritic_b200, checked against the fp64 oracle."""
import numpy as np
import pytest
import torch

import learner_checks as LC
import synth
from oracle import kfac as K
from oracle import learner as OL
from oracle import network as onet

pytestmark = pytest.mark.gpu


def _build(acktr, num_envs, num_steps, seed=0):
    import actorcritic_b200 as ac
    from actorcritic_b200 import kfac, nn, objectives, spaces
    from actorcritic_b200.envs.atari.model import AtariModel
    from actorcritic_b200.kfac_utils import ColdStartPeriodicInvUpdateKfacOpt
    observation_space = spaces.Box(low=0, high=255, shape=(84, 84, 4), dtype=np.uint8)
    action_space = spaces.Discrete(4)
    model = AtariModel(observation_space, action_space, 32 if acktr else 64, random_seed=seed)       # a2c_acktr.py:52-53
    objective = objectives.A2CObjective(model, discount_factor=0.99, entropy_regularization_strength=0.01)   # :57
    global_step = ac.GlobalStep()
    max_step = 1e7 / (num_envs * num_steps)
    if acktr:
        learning_rate = nn.linear_decay(0.25, 0.025, global_step, max_step)                         # :68
        layer_collection = kfac.LayerCollection()                                                    # :235-237
        model.register_layers(layer_collection)
        model.register_predictive_distributions(layer_collection)
        cold = nn.ClipGlobalNormOptimizer(nn.MomentumOptimizer(learning_rate=0.0003, momentum=0.9), clip_norm=0.5)
        optimizer = ColdStartPeriodicInvUpdateKfacOpt(
            num_cold_updates=2, cold_optimizer=cold, invert_every=1, learning_rate=learning_rate, cov_ema_decay=0.99,
            damping=0.01, layer_collection=layer_collection, momentum=0.9, norm_constraint=0.0001,
            cov_devices=["/gpu:0"], inv_devices=["/gpu:0"])
    else:
        learning_rate = nn.linear_decay(0.0007, 0.00007, global_step, max_step)                     # :71
        optimizer = nn.ClipGlobalNormOptimizer(nn.RMSPropOptimizer(learning_rate=learning_rate), clip_norm=0.5)
    optimize_op = objective.optimize_shared(optimizer, baseline_loss_weight=0.5, global_step=global_step)   # :76
    return ac, model, objective, global_step, optimize_op


@pytest.mark.parametrize("acktr", [True, False])
def test_train_step_contract_matches_oracle(acktr):
    e_count, t_count = (4, 5)
    ac, model, objective, global_step, optimize_op = _build(acktr, e_count, t_count)
    params = onet.perturbed_params(4, 32 if acktr else 64, 3)
    model.set_variables(params)
    cfg_o = K.KfacConfig(num_cold_updates=2, invert_every=1, decay_steps=1e7 / 20) if acktr else OL.A2CConfig(decay_steps=1e7 / 20)
    oracle = OL.OracleLearner(params, 4, 32 if acktr else 64, acktr=acktr, cfg=cfg_o)
    with ac.Session() as session:
        for u in range(4):
            batch = synth.rollout(200 + u, e_count, t_count, 4, obs_kind="sparse")
            y_hat, eps = synth.fisher_samples(300 + u, e_count * t_count)
            session.fisher_injection = (torch.from_numpy(y_hat).cuda(), torch.from_numpy(eps).cuda()) if acktr else None
            if u > 0:
                LC.sync_engine_to_oracle(model.engine, oracle)
            # a2c_acktr.py:117-126 (lists of lists are accepted like feed_dict does)
            policy_loss, baseline_loss, entropy, step, _ = session.run(
                [objective.policy_loss, objective.baseline_loss, objective.mean_entropy, global_step, optimize_op],
                feed_dict={model.observations_placeholder: batch["observations"],
                           model.bootstrap_observations_placeholder: batch["bootstrap_observations"],
                           model.actions_placeholder: batch["actions"].tolist(),
                           model.rewards_placeholder: batch["rewards"].tolist(),
                           model.terminals_placeholder: batch["terminals"].tolist()})
            masks = LC.engine_relu_masks(model.engine)       # ReLU derivative at the engine's branch (see learner_checks)
            info = oracle.update(batch, y_hat, eps, masks=masks) if acktr else oracle.update(batch, masks=masks)
            assert step == oracle.global_step
            assert abs(policy_loss - float(info["losses"]["policy_loss"])) <= 1e-4
            assert abs(baseline_loss - float(info["losses"]["baseline_loss"])) <= 1e-4
            assert abs(entropy - float(info["losses"]["mean_entropy"])) <= 1e-4
            got = model.get_variables()
            want = oracle.params_numpy()
            for k in want:
                assert LC.rel_err(got[k], want[k]) <= 2e-4, (u, k)   # variables after the step (step errors ~1e-5 of |step|)
    assert model.engine.config.acktr == acktr and model.engine.config.conv3_filters == (32 if acktr else 64)


def test_sample_and_select_max_actions_shapes():
    ac, model, objective, global_step, optimize_op = _build(True, 6, 3)
    obs = synth.rollout(1, 6, 1, 4, obs_kind="sparse")["observations"]          # [6, 1, 84, 84, 4]
    with ac.Session() as session:
        a = model.sample_actions(obs, session)                                  # model.py:135-151
        m = model.select_max_actions(obs, session)
        assert isinstance(a, list) and len(a) == 6 and all(isinstance(x, int) and 0 <= x < 4 for x in a)
        logits = session.run(model.policy.logits, feed_dict={model.observations_placeholder: obs})
        assert m == np.argmax(logits.reshape(6, 4), axis=1).tolist()
        with pytest.raises(TypeError):
            session.run("not a fetch")
        with pytest.raises(ValueError):
            session.run(optimize_op, feed_dict={model.observations_placeholder: np.zeros((6, 3, 84, 84, 4), np.uint8)})


def test_device_resident_rollout_feeds_the_train_step():
    """MultiEnvAgent over the device-resident synthetic Atari environment: K-PRE fills the rollout buffer, the 6-tuple
    keeps the [environment, step] layout (agents.py:26-45), and it feeds the train step unchanged."""
    from actorcritic_b200 import agents
    from actorcritic_b200.envs.atari.device_env import DeviceAtariMultiEnv
    from oracle import preprocess as OP
    ac, model, objective, global_step, optimize_op = _build(True, 4, 5)
    env = DeviceAtariMultiEnv(4, pool_frames=16, terminal_prob=0.2, seed=1)
    agent = agents.MultiEnvAgent(env, model, num_steps=5)
    with ac.Session() as session:
        obs, act, rew, term, nxt, infos = agent.interact(session)
        assert tuple(obs.shape) == (4, 5, 84, 84, 4) and obs.dtype == torch.uint8 and tuple(nxt.shape) == (4, 84, 84, 4)
        assert tuple(act.shape) == (4, 5) and tuple(rew.shape) == (4, 5) and tuple(term.shape) == (4, 5)
        # the first observation is reset() = 4 copies of the preprocessed first frame (wrappers.py:232-235), bit-exact
        first = OP.preprocess_frame(env.pool[0, 0].cpu().numpy())
        assert np.array_equal(obs[0, 0].cpu().numpy(), np.repeat(first, 4, axis=-1))
        # second observation = push of preprocess(max(frame1, frame2)) (wrappers.py:64-65,224-230)
        second = OP.batched_stack_step(obs[:, 0].cpu().numpy(), env.pool[1].cpu().numpy(), env.pool[2].cpu().numpy(),
                                       env.terminals[0].bool().cpu().numpy())
        assert np.array_equal(obs[:, 1].cpu().numpy(), second)
        loss, _ = session.run([objective.policy_loss, optimize_op],
                              feed_dict={model.observations_placeholder: obs, model.bootstrap_observations_placeholder: nxt,
                                         model.actions_placeholder: act, model.rewards_placeholder: rew,
                                         model.terminals_placeholder: term})
        assert np.isfinite(loss) and global_step.eval() == 2


def test_summaries_written_like_the_reference_example(tmp_path):
    """a2c_acktr.py:80-133: scalars registered under name scopes, merged, fetched WITH the train step, written to an
    event file; the logged values are the step's own losses."""
    from actorcritic_b200 import summary
    from actorcritic_b200.envs.atari.wrappers import EpisodeInfoWrapper
    summary.reset_default_collection()
    ac, model, objective, global_step, optimize_op = _build(True, 4, 5)
    episode_reward_placeholder = summary.placeholder(np.float32, [])
    with summary.name_scope("model"):
        summary.scalar("policy_loss", objective.policy_loss)
        summary.scalar("baseline_loss", objective.baseline_loss)
        summary.scalar("policy_entropy", objective.mean_entropy)
    with summary.name_scope("environment"):
        summary.scalar("episode_reward", episode_reward_placeholder)
    summary_op = summary.merge_all()
    writer = summary.FileWriter(str(tmp_path), None)
    logged = []
    with ac.Session() as session:
        for u in range(3):
            batch = synth.rollout(900 + u, 4, 5, 4, obs_kind="sparse")
            infos = [[{"episode": {"total_reward": 7.0 + u}} if (e, t) == (1, 2) else {} for t in range(5)] for e in range(4)]
            rewards = EpisodeInfoWrapper.get_episode_rewards_from_info_batch(infos)
            mean_reward = np.nan if np.all(np.isnan(rewards)) else np.nanmean(rewards)
            summ, step, _, pl, bl, ent = session.run(
                [summary_op, global_step, optimize_op, objective.policy_loss, objective.baseline_loss, objective.mean_entropy],
                feed_dict={model.observations_placeholder: batch["observations"],
                           model.bootstrap_observations_placeholder: batch["bootstrap_observations"],
                           model.actions_placeholder: batch["actions"], model.rewards_placeholder: batch["rewards"],
                           model.terminals_placeholder: batch["terminals"], episode_reward_placeholder: mean_reward})
            writer.add_summary(summ, step)
            logged.append((int(step), float(pl), float(bl), float(ent), float(mean_reward)))
        nothing, _ = session.run([summary.no_op(), optimize_op],
                                 feed_dict={model.observations_placeholder: batch["observations"],
                                            model.bootstrap_observations_placeholder: batch["bootstrap_observations"],
                                            model.actions_placeholder: batch["actions"],
                                            model.rewards_placeholder: batch["rewards"],
                                            model.terminals_placeholder: batch["terminals"]})
        assert nothing is None
    writer.close()
    events = [e for e in summary.read_events(writer.path) if e["scalars"]]
    assert len(events) == 3
    for ev, (step, pl, bl, ent, rew) in zip(events, logged):
        assert ev["step"] == step
        assert ev["scalars"] == {"model/policy_loss": np.float32(pl), "model/baseline_loss": np.float32(bl),
                                 "model/policy_entropy": np.float32(ent), "environment/episode_reward": np.float32(rew)}
    summary.reset_default_collection()
