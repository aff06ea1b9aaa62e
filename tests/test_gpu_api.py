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


# ------------------------------------------------------------------------------------------------ nn.* layer ops
def test_nn_layer_ops_match_the_oracle_network():
    """nn.conv2d / nn.flatten / nn.fully_connected (nn.py:37-52,88-126) on device tensors reproduce the oracle's Nature-CNN
    forward (oracle/network.py, pinned to the reference's graph code by tests/golden/network.npz) layer by layer; the
    implicit-GEMM route (`acx_conv`) agrees with the im2col route on the two layers it covers."""
    from actorcritic_b200 import nn
    params = onet.perturbed_params(4, 32, 5)
    obs = synth.rollout(31, 3, 2, 4, obs_kind="sparse")["observations"].reshape(6, 84, 84, 4)
    fwd = onet.forward(onet.to_torch(params), obs)
    x = torch.from_numpy(obs).cuda().float() / 255.0                       # envs/atari/model.py:93
    acts = {}
    for name, stride in (("conv1", 4), ("conv2", 2), ("conv3", 1)):
        y = nn.conv2d(x, (params[name + "/weights"], params[name + "/bias"]), stride, "VALID")
        if name != "conv1":
            y2 = nn.conv2d(x, (params[name + "/weights"], params[name + "/bias"]), stride, "VALID", impl="implicit")
            assert LC.rel_err(y2.cpu().numpy(), y.cpu().numpy()) <= 1e-5      # three bf16 output planes re-summed
        want = fwd[name]["pre"].numpy().reshape(tuple(y.shape))
        assert LC.rel_err(y.cpu().numpy(), want) <= 1e-5, name
        x = torch.relu(y)
        acts[name] = x
    flat = nn.flatten(x)
    assert tuple(flat.shape) == (6, 49 * 32)
    np.testing.assert_array_equal(flat.cpu().numpy(), acts["conv3"].cpu().numpy().reshape(6, -1))    # (h, w, c) order
    h = torch.relu(nn.fully_connected(flat, (params["fc4/weights"], params["fc4/bias"])))
    logits = nn.fully_connected(h, (params["fc_policy/weights"], params["fc_policy/bias"]))
    value = nn.fully_connected(h, (params["fc_baseline/weights"], params["fc_baseline/bias"]))
    assert LC.rel_err(logits.cpu().numpy(), fwd["logits"].numpy()) <= 1e-4      # five layers deep
    assert LC.rel_err(value.cpu().numpy()[:, 0], fwd["value"].numpy()) <= 1e-4


def test_nn_conv2d_same_padding_and_odd_geometry():
    """A geometry outside the Nature-CNN (5x5 / 3, 3 -> 10 channels, SAME and VALID) against torch's convolution in fp64."""
    from actorcritic_b200 import nn
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 17, 17, 3)).astype(np.float32)
    w = (0.2 * rng.standard_normal((5, 5, 3, 10))).astype(np.float32)
    b = rng.standard_normal(10).astype(np.float32)
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).double().permute(3, 2, 0, 1)
    for padding, pad in (("VALID", 0), ("SAME", None)):
        got = nn.conv2d(torch.from_numpy(x).cuda(), (w, b), 3, padding).cpu().numpy()
        if pad is None:      # TF SAME for 17 / stride 3 / k 5: out 6, total padding 3 -> (1, 2)
            want = torch.nn.functional.conv2d(torch.nn.functional.pad(xt, (1, 2, 1, 2)), wt, torch.from_numpy(b).double(), stride=3)
        else:
            want = torch.nn.functional.conv2d(xt, wt, torch.from_numpy(b).double(), stride=3)
        want = want.permute(0, 2, 3, 1).numpy()
        assert got.shape == want.shape and LC.rel_err(got, want) <= 1e-5, padding


def test_non_atari_model_built_from_nn_layers():
    """An ActorCriticModel subclass for a Box observation space assembled from nn.* like the reference assembles AtariModel
    (model.py:107-133, envs/atari/model.py:173-217): sample_actions / select_max_actions through Session.run."""
    import actorcritic_b200 as ac
    from actorcritic_b200 import nn, spaces
    from actorcritic_b200.baselines import StateValueFunction
    from actorcritic_b200.model import ActorCriticModel
    from actorcritic_b200.policies import SoftmaxPolicy

    class VectorModel(ActorCriticModel):
        def __init__(self, observation_space, action_space, seed=0):
            super().__init__(observation_space, action_space)
            rng = np.random.default_rng(seed)
            self.random_seed = seed
            self.fc1 = nn.fully_connected_params(observation_space.shape[0], 24, rng=rng, gain=2 ** 0.5)
            self.fc_policy = nn.fully_connected_params(24, action_space.n, rng=rng, gain=1.0)
            self.fc_baseline = nn.fully_connected_params(24, 1, rng=rng, gain=1.0)
            self._policy = SoftmaxPolicy(self, action_space.n)
            self._baseline = StateValueFunction(self)

        def _forward_device(self, observations):
            h = torch.relu(nn.fully_connected(nn.flatten(observations.float()), self.fc1))
            return nn.fully_connected(h, self.fc_policy), nn.fully_connected(h, self.fc_baseline)[:, 0]

    observation_space = spaces.Box(low=-1.0, high=1.0, shape=(11,), dtype=np.float32)
    model = VectorModel(observation_space, spaces.Discrete(5), seed=2)
    assert model.observations_placeholder.dtype == np.float32 and model.observations_placeholder.shape == (None, None, 11)
    obs = np.random.default_rng(1).uniform(-1, 1, (7, 1, 11)).astype(np.float32)
    w1, b1 = model.fc1
    h = np.maximum(obs.reshape(7, 11).astype(np.float64) @ w1.astype(np.float64) + b1, 0.0)
    want_logits = h @ model.fc_policy[0].astype(np.float64) + model.fc_policy[1]
    with ac.Session() as session:
        greedy = model.select_max_actions(obs, session)
        sampled = model.sample_actions(obs, session)
        logits, value = session.run([model.policy.logits, model.baseline.value],
                                    feed_dict={model.observations_placeholder: obs})
    assert LC.rel_err(logits.reshape(7, 5), want_logits) <= 1e-5
    assert greedy == want_logits.argmax(1).tolist()
    assert len(sampled) == 7 and all(0 <= a < 5 for a in sampled)
    assert value.shape == (7, 1)
    with pytest.raises(NotImplementedError):
        model.register_layers(None)


def test_box_action_space_placeholders():
    """model.py:179-183: a Box action space gives a placeholder of the space's dtype and shape (f4)."""
    from actorcritic_b200 import spaces
    from actorcritic_b200.model import ActorCriticModel

    class M(ActorCriticModel):
        pass
    m = M(spaces.Box(low=0, high=255, shape=(84, 84, 4), dtype=np.uint8), spaces.Box(low=-2.0, high=2.0, shape=(3,), dtype=np.float32))
    assert m.actions_placeholder.dtype == np.float32 and m.actions_placeholder.shape == (None, None, 3)
    assert m.observations_placeholder.dtype == np.uint8 and m.bootstrap_observations_placeholder.shape == (None, 84, 84, 4)


# ------------------------------------------------------------------------------------------------ optimize_separate (f4)
def test_optimize_separate_matches_oracle():
    """objectives.py:31-54: policy loss and baseline loss minimised by two optimizers (two backward passes from the same
    parameters, separate slots), against the oracle's gradients of each loss and the TF-1 RMSProp / Momentum recurrences."""
    import actorcritic_b200 as ac
    from actorcritic_b200 import nn, objectives, spaces
    from actorcritic_b200.envs.atari.model import AtariModel
    e_count, t_count, c3 = 4, 5, 64
    model = AtariModel(spaces.Box(low=0, high=255, shape=(84, 84, 4), dtype=np.uint8), spaces.Discrete(4), c3, random_seed=0)
    objective = objectives.A2CObjective(model, discount_factor=0.99, entropy_regularization_strength=0.01)
    with pytest.raises(TypeError):      # the reference's default None kwargs: `**None`
        objective.optimize_separate(nn.RMSPropOptimizer(1e-3), nn.RMSPropOptimizer(1e-3))
    global_step = ac.GlobalStep()
    policy_opt = nn.ClipGlobalNormOptimizer(nn.RMSPropOptimizer(learning_rate=7e-4), clip_norm=0.5)
    baseline_opt = nn.MomentumOptimizer(learning_rate=3e-4, momentum=0.9)
    op = objective.optimize_separate(policy_opt, baseline_opt, policy_kwargs=dict(global_step=global_step), baseline_kwargs={})
    params = onet.perturbed_params(4, c3, 3)
    model.set_variables(params)
    o = OL.OracleLearner(params, 4, c3, acktr=False)
    ms = {l: torch.ones_like(onet.join_vmat(l, o.params)) for l in onet.LAYERS}
    mom = {l: torch.zeros_like(onet.join_vmat(l, o.params)) for l in onet.LAYERS}
    with ac.Session() as session:
        for u in range(3):
            batch = synth.rollout(400 + u, e_count, t_count, 4, obs_kind="sparse")
            if u > 0:      # compare each update from identical parameters
                for l in onet.LAYERS:
                    model.engine.layer_matrix("params", l).copy_(onet.join_vmat(l, o.params).float())
                model.engine.refresh_derived()
            pl, bl, _ = session.run([objective.policy_loss, objective.baseline_loss, op], feed_dict={
                model.observations_placeholder: batch["observations"],
                model.bootstrap_observations_placeholder: batch["bootstrap_observations"],
                model.actions_placeholder: batch["actions"], model.rewards_placeholder: batch["rewards"],
                model.terminals_placeholder: batch["terminals"]})
            e = model.engine
            masks = LC.engine_relu_masks(e)
            info = o.compute(batch, need_fisher=False, masks=masks)
            n = e_count * t_count
            actions = batch["actions"].reshape(n)
            dz, _ = onet.output_grads(info["fwd"]["logits"], info["fwd"]["value"], actions, info["targets"], 0.01, 0.0)
            _, dv = onet.output_grads(info["fwd"]["logits"], info["fwd"]["value"], actions, info["targets"], 0.01, 1.0)
            g_pol, _ = onet.backward(o.params, info["fwd"], dz, torch.zeros_like(dv), masks)
            g_base, _ = onet.backward(o.params, info["fwd"], torch.zeros_like(dz), dv, masks)
            clipped, _ = K.clip_by_global_norm(g_pol, 0.5)
            before = LC.oracle_flat_params(o)
            for l in onet.LAYERS:
                ms[l] = 0.9 * ms[l] + 0.1 * clipped[l] * clipped[l]
                mom[l] = 0.9 * mom[l] + g_base[l]
                new = onet.join_vmat(l, o.params) - 7e-4 * clipped[l] / torch.sqrt(ms[l] + 1e-10) - 3e-4 * mom[l]
                w, b = onet.split_vmat(l, new, o.params)
                o.params[l + "/weights"], o.params[l + "/bias"] = w, b
            assert abs(float(pl) - float(info["losses"]["policy_loss"])) <= 2e-4 * max(1.0, abs(float(info["losses"]["policy_loss"])))
            assert abs(float(bl) - float(info["losses"]["baseline_loss"])) <= 2e-4 * max(1.0, abs(float(info["losses"]["baseline_loss"])))
            got_step = e.get_params_flat().astype(np.float64) - before
            want_step = LC.oracle_flat_params(o) - before
            # the step itself: fp32 parameter storage leaves ~1e-8 |theta| / |step| of noise (as in test_a2c_rmsprop_matches_oracle)
            assert LC.rel_err(got_step, want_step) <= 1e-2, u
            assert global_step.eval() == u + 1          # only the policy optimizer was handed the global step


def test_graph_replayed_rollout_equals_stepwise_rollout():
    """MultiEnvAgent.interact on a device environment whose rollouts repeat (pool_frames = T) replays the whole rollout as
    one CUDA graph from its third call on: observations, actions (Philox stream continues through the device-resident
    call counter), rewards, terminals and next observations are identical to stepping one by one."""
    import actorcritic_b200 as ac
    from actorcritic_b200 import agents, engine as eng
    from actorcritic_b200.envs.atari.device_env import DeviceAtariMultiEnv

    def run(use_graphs):
        e = eng.Engine(eng.EngineConfig(num_envs=4, num_steps=5, seed=11, use_graphs=use_graphs))
        e.set_params(onet.perturbed_params(4, 32, 2))

        class M:
            engine = e
        env = DeviceAtariMultiEnv(4, pool_frames=5, terminal_prob=0.2, seed=3)
        agent = agents.MultiEnvAgent(env, M(), 5)
        outs = []
        for _ in range(5):
            o, a, r, t, nxt, infos = agent.interact(None)
            torch.cuda.synchronize()
            outs.append((o.cpu().numpy().copy(), a.cpu().numpy().copy(), r.cpu().numpy().copy(), t.cpu().numpy().copy(),
                         nxt.cpu().numpy().copy()))
            assert len(infos) == 4 and len(infos[0]) == 5
        return outs, getattr(agent, "_graph_state", 0)
    eager, st0 = run(False)
    graph, st1 = run(True)
    assert st0 == 0 and st1 == 2
    for k, (x, y) in enumerate(zip(eager, graph)):
        for i, name in enumerate(("observations", "actions", "rewards", "terminals", "next_observations")):
            assert np.array_equal(x[i], y[i]), (k, name)
    assert len({tuple(x[1].ravel()) for x in eager}) > 1      # the sampled actions do change from rollout to rollout
