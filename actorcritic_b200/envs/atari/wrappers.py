"""The array-processing subset of actorcritic/envs/atari/wrappers.py, batched over environments and run by the
K-PRE kernel (preprocess.cu): 2-frame max (wrappers.py:64-65), RGB->gray + 84x84 INTER_AREA resize (:30-33), frame
stack push / zero-on-terminal / reset-to-4-copies (:224-235) in the order MultiEnv's auto-reset imposes
(multi_env.py:127-132).  Bit-exact with the reference's cv2 + NumPy path (tests/test_gpu_preprocess.py).

`BatchedAtariPreprocessor` is the batched form; the per-environment classes keep the reference's names and call
the same kernel with one environment."""
import numpy as np
import torch

from ... import ops
from ... import spaces


class BatchedAtariPreprocessor:
    """Frame stacks of E environments on the device."""

    def __init__(self, num_envs, device=None, num_stacked_frames=4):
        if num_stacked_frames != 4:
            raise NotImplementedError("the K-PRE kernel packs exactly 4 stacked frames into one 32-bit word per pixel")
        self.num_envs = num_envs
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.stacks = torch.zeros((num_envs, 84, 84, 4), dtype=torch.uint8, device=self.device)
        self.prev_terminal = torch.zeros(num_envs, dtype=torch.uint8, device=self.device)

    def reset(self, raw_frames):
        """MultiEnv.reset: stack = 4 copies of preprocess(frame) (wrappers.py:232-235)."""
        ops.preprocess_reset(raw_frames, out=self.stacks, out_env_stride=28224)
        self.prev_terminal.zero_()
        return self.stacks

    def step(self, raw_a, raw_b, terminal, reset_raw=None, out=None, out_env_stride=None, keep_terminal_view=False):
        """One MultiEnv.step: raw_a/raw_b = the last two emulator frames of each frameskip window, `terminal` = this
        step's terminal flags; environments whose PREVIOUS step was terminal are first reset from reset_raw.
        `out` may be a slice [:, t] of a batch-major rollout buffer."""
        reset_mask = self.prev_terminal if reset_raw is not None else None
        ops.preprocess_stack(raw_a, raw_b, self.stacks, terminal=terminal, reset_mask=reset_mask, reset_raw=reset_raw,
                             out=self.stacks, out_env_stride=28224)
        if out is not None:
            out.copy_(self.stacks)
        if terminal is None:
            self.prev_terminal = torch.zeros_like(self.prev_terminal)
        else:       # keep_terminal_view: the caller guarantees that `terminal` is not overwritten before the next step
            self.prev_terminal = terminal if keep_terminal_view else terminal.clone()
        return self.stacks


class AtariPreprocessFrameWrapper:
    """wrappers.py:16-33 for one environment: observation(frame) -> uint8 [84,84,1]."""

    def __init__(self, env):
        self.env = env
        self.observation_space = spaces.Box(low=0, high=255, shape=(84, 84, 1), dtype=np.uint8)
        self.action_space = getattr(env, "action_space", None)

    def observation(self, frame):
        raw = torch.from_numpy(np.ascontiguousarray(frame, dtype=np.uint8)[None]).cuda()
        return ops.preprocess_reset(raw)[0, :, :, 3:4].cpu().numpy()

    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        observation, reward, terminal, info = self.env.step(action)
        return self.observation(observation), reward, terminal, info


class AtariFrameskipWrapper:
    """wrappers.py:36-70 for one environment: repeats the action `frameskip` times, sums the rewards, stops at a terminal
    step, and returns the byte-wise max of the last two frames (acx_frame_max_u8) - or the only frame when the first
    sub-step was terminal."""

    def __init__(self, env, frameskip):
        self.env = env
        self._frameskip = frameskip
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    def step(self, action):
        older = newest = None            # only the last two frames of the window matter
        reward_sum, done, info = 0.0, False, None
        for _ in range(self._frameskip):
            frame, reward, done, info = self.env.step(action)
            older, newest = newest, frame
            reward_sum += reward
            if done:
                break
        if older is None:                # the first sub-step ended the episode: a one-frame window (wrappers.py:66-67)
            return newest, reward_sum, done, info
        pair = torch.from_numpy(np.ascontiguousarray(np.stack([older, newest]), dtype=np.uint8)).cuda()
        return ops.frame_max(pair[0], pair[1]).cpu().numpy(), reward_sum, done, info

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


class FrameStackWrapper:
    """wrappers.py:201-235 for one environment that already yields [84,84,1] frames (host arrays); the stack lives on the
    device and is pushed by acx_framestack_push_u8."""

    def __init__(self, env, num_stacked_frames):
        if num_stacked_frames != 4:
            raise NotImplementedError("4 stacked frames (a2c_acktr.py:171)")
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = spaces.Box(low=0, high=255, shape=(84, 84, 4), dtype=np.uint8)
        self._stack = torch.zeros((1, 84, 84, 4), dtype=torch.uint8, device="cuda")

    def _push(self, frame, mode):
        f = torch.from_numpy(np.ascontiguousarray(frame, np.uint8).reshape(1, 84, 84)).cuda()
        m = torch.tensor([mode], dtype=torch.uint8, device="cuda")
        ops.framestack_push(f, self._stack, mode=m, out=self._stack)
        return self._stack[0].cpu().numpy()

    def step(self, action):
        frame, reward, terminal, info = self.env.step(action)
        return self._push(frame, 1 if terminal else 0), reward, terminal, info       # wrappers.py:226-229

    def reset(self, **kwargs):
        return self._push(self.env.reset(**kwargs), 2)                                 # wrappers.py:234


class EpisodeInfoWrapper:
    """wrappers.py:263-323: accumulates the rewards of an episode and, on its terminal step, reports them as
    info['episode']['total_reward'] (host-side bookkeeping; stacks on any environment, also on those a RawFrameMultiEnv
    hosts)."""

    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)
        self.total_reward = 0.0

    def step(self, action):
        observation, reward, terminal, info = self.env.step(action)
        self.total_reward += reward
        if terminal:
            info = dict(info) if info is not None else {}
            info["episode"] = {"total_reward": self.total_reward}
            self.total_reward = 0.0
        return observation, reward, terminal, info

    def reset(self, **kwargs):
        self.total_reward = 0.0
        return self.env.reset(**kwargs)

    @staticmethod
    def get_episode_rewards_from_info_batch(infos):
        """Batch-major [environment][step] infos (as `Agent.interact` returns them) -> float32 array of the same shape
        with the episode reward where an episode ended and NaN elsewhere (wrappers.py:296-323); the caller takes
        `np.nanmean` of it for the 'episode_reward' summary (a2c_acktr.py:112-114)."""
        environments = len(infos)
        steps = len(infos[0]) if environments else 0
        rewards = np.full((environments, steps), np.nan, np.float32)
        for e, row in enumerate(infos):
            for t, info in enumerate(row):
                episode = info.get("episode") if info else None
                if episode is not None:
                    rewards[e, t] = episode["total_reward"]
        return rewards
