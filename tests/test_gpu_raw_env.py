"""RawFrameMultiEnv: environments that yield raw 210x160x3 frames, frame work on the GPU.  Checked
bit-for-bit against the reference's wrapper chain restated on the host: AtariFrameskipWrapper (wrappers.py:52-67) ->
AtariPreprocessFrameWrapper (:30-33) -> FrameStackWrapper (:224-235) under MultiEnv's _AutoResetWrapper
(multi_env.py:127-137)."""
import numpy as np
import pytest
import torch

from oracle import preprocess as P

pytestmark = pytest.mark.gpu


class FakeAtari:
    """Deterministic emulator stand-in: frame t of episode k is a seeded random image; the episode ends after a scripted
    number of emulator steps; reward = (t % 3) - 1."""

    def __init__(self, seed, episode_lengths, num_actions=4):
        from actorcritic_b200 import spaces
        self.action_space = spaces.Discrete(num_actions)
        self.seed, self.lengths = seed, list(episode_lengths)
        self.episode, self.t = -1, 0
        self.closed = False

    def _frame(self):
        rng = np.random.default_rng((self.seed, self.episode, self.t))
        kind = (self.episode + self.t) % 3
        if kind == 0:
            return rng.integers(0, 256, (210, 160, 3), dtype=np.uint8)
        if kind == 1:                                   # flat colour blocks, Atari-like
            f = np.zeros((210, 160, 3), np.uint8)
            f[rng.integers(0, 200):, rng.integers(0, 150):] = rng.integers(0, 256, 3, dtype=np.uint8)
            return f
        return np.full((210, 160, 3), rng.integers(0, 256), np.uint8)

    def reset(self):
        self.episode += 1
        self.t = 0
        return self._frame()

    def step(self, action):
        self.t += 1
        terminal = self.t >= self.lengths[self.episode % len(self.lengths)]
        return self._frame(), float(self.t % 3 - 1), terminal, {"t": self.t, "episode": self.episode, "action": action}

    def close(self):
        self.closed = True


class HostChain:
    """The reference's wrapper chain for one environment, array work by the oracle."""

    def __init__(self, env, frameskip):
        self.env, self.frameskip = env, frameskip
        self.stack = P.FrameStack(4)
        self.terminated = False

    def reset(self):
        self.terminated = False
        return self.stack.reset(P.preprocess_frame(self.env.reset())).copy()

    def step(self, action):
        if self.terminated:                             # multi_env.py:128-129
            self.stack.reset(P.preprocess_frame(self.env.reset()))
        frames, total, terminal, info = [], 0.0, False, None
        for _ in range(self.frameskip):                 # wrappers.py:52-62
            f, r, terminal, info = self.env.step(action)
            frames.append(f)
            total += r
            if terminal:
                break
        frame = np.amax((frames[-2], frames[-1]), axis=0) if len(frames) >= 2 else frames[0]
        obs = self.stack.step(P.preprocess_frame(frame), terminal).copy()
        self.terminated = terminal
        return obs, total, terminal, info


def _lengths(i):
    return [(5, 9, 1, 13), (4, 4, 4), (17,), (1, 2, 3), (8, 1, 8)][i % 5]


@pytest.mark.parametrize("frameskip", [4, 1])
def test_raw_frame_multi_env_matches_the_reference_wrapper_chain(frameskip):
    from actorcritic_b200.envs.atari.raw_env import RawFrameMultiEnv
    e_count, steps = 5, 14
    env = RawFrameMultiEnv([FakeAtari(10 + i, _lengths(i)) for i in range(e_count)], frameskip=frameskip)
    chains = [HostChain(FakeAtari(10 + i, _lengths(i)), frameskip) for i in range(e_count)]
    got = env.reset()
    want = np.stack([c.reset() for c in chains])
    assert got.dtype == torch.uint8 and tuple(got.shape) == (e_count, 84, 84, 4)
    assert np.array_equal(got.cpu().numpy(), want)
    rng = np.random.default_rng(0)
    saw_terminal = saw_single_frame = False
    for t in range(steps):
        actions = rng.integers(0, 4, e_count).tolist()
        obs, rew, term = env.step_device(torch.tensor(actions, dtype=torch.int32, device="cuda"))
        ref = [c.step(a) for c, a in zip(chains, actions)]
        assert np.array_equal(obs.cpu().numpy(), np.stack([r[0] for r in ref])), "step %d" % t
        assert rew.cpu().tolist() == [r[1] for r in ref]
        assert term.bool().cpu().tolist() == [r[2] for r in ref]
        assert env.last_infos == [r[3] for r in ref]
        saw_terminal |= any(r[2] for r in ref)
        saw_single_frame |= any(r[2] and r[3]["t"] == 1 for r in ref)
    assert saw_terminal and (frameskip == 1 or saw_single_frame)     # auto-reset and the one-frame window were exercised
    # host-list form of MultiEnv.step (multi_env.py:59-81)
    o, r, d, infos = env.step([0] * e_count)
    ref = [c.step(0) for c in chains]
    assert isinstance(o, list) and np.array_equal(np.stack(o), np.stack([x[0] for x in ref])) and d == [x[2] for x in ref]
    env.close()
    assert all(e.closed for e in env.envs)


def test_agent_rollout_over_raw_frame_envs_feeds_the_train_step():
    from actorcritic_b200 import agents
    from actorcritic_b200.envs.atari.raw_env import RawFrameMultiEnv
    from test_gpu_api import _build
    ac, model, objective, global_step, optimize_op = _build(True, 4, 5)
    env = RawFrameMultiEnv([FakeAtari(30 + i, _lengths(i)) for i in range(4)], frameskip=4)
    agent = agents.MultiEnvAgent(env, model, num_steps=5)
    with ac.Session() as session:
        obs, act, rew, term, nxt, infos = agent.interact(session)
        assert tuple(obs.shape) == (4, 5, 84, 84, 4) and tuple(nxt.shape) == (4, 84, 84, 4)
        assert len(infos) == 4 and len(infos[0]) == 5 and "episode" in infos[0][0]      # [environment][step] (agents.py:26-45)
        loss, _ = session.run([objective.policy_loss, optimize_op],
                              feed_dict={model.observations_placeholder: obs, model.bootstrap_observations_placeholder: nxt,
                                         model.actions_placeholder: act, model.rewards_placeholder: rew,
                                         model.terminals_placeholder: term})
        assert np.isfinite(loss)
        obs2, *_ = agent.interact(session)
        assert torch.equal(obs2[:, 0], nxt)             # the agent carries the last observations over (agents.py:155,219)
    env.close()


def test_reward_clipping_and_step_hook_act_on_agent_steps():
    """The reference clips rewards ABOVE the frameskip (a2c_acktr.py:190-208): the clip applies to the 4-frame sum, not to
    each emulator frame.  `clip_rewards` / `step_hook` of RawFrameMultiEnv act there."""
    from actorcritic_b200.envs.atari.raw_env import RawFrameMultiEnv

    class Rewarding(FakeAtari):
        def step(self, action):
            f, _, terminal, info = super().step(action)
            return f, 0.75, terminal, info               # 4 frames sum to 3.0: clipped AFTER the sum -> 1.0 (per frame it would be 3.0)

    seen = []

    def hook(i, reward, terminal, info):
        seen.append((i, reward, terminal))
        return reward * 2.0, terminal, dict(info, hooked=True)

    envs = [Rewarding(5 + i, (50,)) for i in range(3)]
    multi = RawFrameMultiEnv(envs, frameskip=4, clip_rewards=True, step_hook=hook)
    multi.reset()
    _, rew, term = multi.step_device(torch.zeros(3, dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert rew.cpu().tolist() == [2.0, 2.0, 2.0]        # clip(3.0) = 1.0, then the hook doubles it
    assert sorted(s[0] for s in seen) == [0, 1, 2] and all(s[1] == 1.0 for s in seen)
    assert all(info.get("hooked") for info in multi.last_infos)
    multi.close()
