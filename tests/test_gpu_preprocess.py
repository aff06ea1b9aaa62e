"""G1/G2 on the GPU: K-PRE (through the C ABI) is bit-exact against the oracle and the reference golden."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import preprocess as P

pytestmark = pytest.mark.gpu


def _ops():
    from actorcritic_b200 import ops
    return ops


@pytest.mark.parametrize("kind", ["uniform", "palette", "blocky", "binary"])
def test_reset_matches_oracle_bit_exact(kind):
    ops = _ops()
    frames = synth.raw_frames(3, 9, kind)
    got = ops.preprocess_reset(torch.from_numpy(frames).cuda()).cpu().numpy()
    want = np.repeat(np.stack([P.preprocess_frame(f) for f in frames]), 4, axis=-1)
    assert np.array_equal(got, want)


def test_golden_preprocess_and_max(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    frames = synth.raw_frames(int(g["seed"]), int(g["count"]), "mixed")
    got = ops.preprocess_reset(torch.from_numpy(frames).cuda()).cpu().numpy()
    assert np.array_equal(got[..., 3:], g["observation"])
    # max of (f2,f3) and (f4,f5) through the step kernel: newest channel = preprocess(max)
    a = torch.from_numpy(frames[[2, 4]]).cuda()
    b = torch.from_numpy(frames[[3, 5]]).cuda()
    out = ops.preprocess_stack(a, b, torch.zeros((2, 84, 84, 4), dtype=torch.uint8, device="cuda")).cpu().numpy()
    assert np.array_equal(out[0, ..., 3], P.preprocess_frame(g["skip_max_23"])[..., 0])
    assert np.array_equal(out[1, ..., 3], P.preprocess_frame(g["skip_max_45"])[..., 0])
    assert not out[..., :3].any()


def test_framestack_sequence_matches_reference_golden(golden_dir):
    """Drive the kernel through the scripted terminal/reset sequence the reference's FrameStackWrapper +
    _AutoResetWrapper were driven through (make_golden.py) and compare every observation."""
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "framestack.npz"))
    num_steps = g["observations"].shape[0]
    frames = [synth.raw_frames(int(g["seeds"][e]), int(g["frames_per_env"]), "mixed") for e in range(2)]
    traces = [list(map(tuple, g["trace_env%d" % e])) for e in range(2)]
    pos = [0, 0]
    first = np.stack([frames[e][traces[e][0][0]] for e in range(2)])
    pos = [1, 1]
    stacks = ops.preprocess_reset(torch.from_numpy(first).cuda())
    assert np.array_equal(stacks.cpu().numpy(), g["reset_observation"])
    for t in range(num_steps):
        raw_a = np.zeros((2, 210, 160, 3), np.uint8)
        raw_b = np.zeros_like(raw_a)
        reset_raw = np.zeros_like(raw_a)
        reset_mask = np.zeros(2, np.uint8)
        for e in range(2):
            a, b, is_reset = traces[e][pos[e]]
            if is_reset:
                reset_mask[e] = 1
                reset_raw[e] = frames[e][a]
                pos[e] += 1
                a, b, is_reset = traces[e][pos[e]]
            raw_a[e], raw_b[e] = frames[e][a], frames[e][b]
            pos[e] += 1
        term = torch.from_numpy(g["terminals"][t].astype(np.uint8)).cuda()
        stacks = ops.preprocess_stack(torch.from_numpy(raw_a).cuda(), torch.from_numpy(raw_b).cuda(), stacks,
                                      terminal=term, reset_mask=torch.from_numpy(reset_mask).cuda(),
                                      reset_raw=torch.from_numpy(reset_raw).cuda())
        assert np.array_equal(stacks.cpu().numpy(), g["observations"][t]), "step %d" % t


@pytest.mark.parametrize("num_envs", [1, 3, 64, 257, 700])
def test_random_steps_against_oracle(num_envs):
    """257 environments run the one-band-per-CTA kernel, 700 the persistent TMA-pipelined one (>= 8 bands per CTA)."""
    ops = _ops()
    rng = np.random.default_rng(num_envs)
    kinds = ["uniform", "palette", "blocky", "binary"]
    stacks = rng.integers(0, 256, (num_envs, 84, 84, 4), dtype=np.uint8)
    pool = np.concatenate([synth.raw_frames(50 + i, 6, k) for i, k in enumerate(kinds)])
    ia, ib, ir = (rng.integers(0, len(pool), num_envs) for _ in range(3))
    term = (rng.random(num_envs) < 0.3)
    reset = (rng.random(num_envs) < 0.3)
    want = P.batched_stack_step(stacks, pool[ia], pool[ib], term, reset, pool[ir])
    got = ops.preprocess_stack(torch.from_numpy(pool[ia]).cuda(), torch.from_numpy(pool[ib]).cuda(),
                               torch.from_numpy(stacks).cuda(), terminal=torch.from_numpy(term.astype(np.uint8)).cuda(),
                               reset_mask=torch.from_numpy(reset.astype(np.uint8)).cuda(),
                               reset_raw=torch.from_numpy(pool[ir]).cuda())
    assert np.array_equal(got.cpu().numpy(), want)


def test_writes_into_rollout_buffer_and_in_place():
    ops = _ops()
    e_count, t_count = 5, 3
    frames = synth.raw_frames(9, 2 * e_count, "mixed")
    a = torch.from_numpy(frames[:e_count]).cuda()
    b = torch.from_numpy(frames[e_count:]).cuda()
    prev = torch.randint(0, 256, (e_count, 84, 84, 4), dtype=torch.uint8, device="cuda")
    want = ops.preprocess_stack(a, b, prev)
    rollout = torch.zeros((e_count, t_count, 84, 84, 4), dtype=torch.uint8, device="cuda")
    ops.preprocess_stack(a, b, prev, out=rollout[:, 1], out_env_stride=rollout.stride(0))
    assert torch.equal(rollout[:, 1], want) and not rollout[:, 0].any() and not rollout[:, 2].any()
    inplace = prev.clone()
    ops.preprocess_stack(a, b, inplace, out=inplace)
    assert torch.equal(inplace, want)


def test_empty_batch_is_a_noop():
    ops = _ops()
    z = torch.zeros((0, 210, 160, 3), dtype=torch.uint8, device="cuda")
    out = ops.preprocess_stack(z, z, torch.zeros((0, 84, 84, 4), dtype=torch.uint8, device="cuda"))
    assert out.shape == (0, 84, 84, 4)


def test_idempotent_channel_shift_property_large():
    """Size-independent property at the sweep's largest size: after 4 pushes of the same frame pair the
    stack equals the reset stack of that frame."""
    ops = _ops()
    e_count = 1024
    pool = synth.raw_frames(77, 8, "mixed")
    idx = np.arange(e_count) % 8
    a = torch.from_numpy(pool[idx]).cuda()
    stacks = torch.randint(0, 256, (e_count, 84, 84, 4), dtype=torch.uint8, device="cuda")
    for _ in range(4):
        stacks = ops.preprocess_stack(a, a, stacks)
    assert torch.equal(stacks, ops.preprocess_reset(a))


def test_standalone_frame_max_and_framestack_wrappers_match_the_reference_lines():
    """The per-environment wrapper classes of the reference's API (wrappers.py:36-70, 201-235) on the stand-alone C-ABI
    entry points, against the reference's NumPy lines restated by the oracle."""
    from actorcritic_b200.envs.atari import wrappers as W
    ops = _ops()
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (3, 210, 160, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (3, 210, 160, 3), dtype=np.uint8)
    assert np.array_equal(ops.frame_max(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy(), np.maximum(a, b))
    odd = rng.integers(0, 256, 1003, dtype=np.uint8)            # ragged tail (not a multiple of 16 bytes)
    odd2 = rng.integers(0, 256, 1003, dtype=np.uint8)
    assert np.array_equal(ops.frame_max(torch.from_numpy(odd).cuda(), torch.from_numpy(odd2).cuda()).cpu().numpy(),
                          np.maximum(odd, odd2))

    class Scripted:
        def __init__(self, frames, terminals):
            self.frames, self.terminals, self.i = frames, terminals, 0
            self.action_space = None

        def reset(self):
            self.i += 1
            return self.frames[self.i - 1]

        def step(self, action):
            self.i += 1
            return self.frames[self.i - 1], 1.0, self.terminals[self.i - 1], {"i": self.i}

    # AtariFrameskipWrapper: max of the last two of 4 frames; a terminal first sub-step returns that single frame
    raw = [rng.integers(0, 256, (210, 160, 3), dtype=np.uint8) for _ in range(12)]
    term = [False] * 12
    term[6] = True                                               # third sub-step of the second window
    term[7] = True                                               # first sub-step of the third window
    env = W.AtariFrameskipWrapper(Scripted(raw, term), 4)
    assert np.array_equal(env.reset(), raw[0])
    obs, rew, done, _ = env.step(0)
    assert np.array_equal(obs, np.maximum(raw[3], raw[4])) and rew == 4.0 and not done
    obs, rew, done, _ = env.step(0)
    assert np.array_equal(obs, np.maximum(raw[5], raw[6])) and rew == 2.0 and done
    obs, rew, done, _ = env.step(0)
    assert np.array_equal(obs, raw[7]) and rew == 1.0 and done

    # FrameStackWrapper: reset = 4 copies, step = roll / zero on terminal / newest last
    gray = [rng.integers(0, 256, (84, 84, 1), dtype=np.uint8) for _ in range(6)]
    fs_term = [False, False, False, True, False, False]
    fs = W.FrameStackWrapper(Scripted(gray, fs_term), 4)
    ref = P.FrameStack(4)
    assert np.array_equal(fs.reset(), ref.reset(gray[0]))
    for i in range(1, 6):
        obs, _, done, _ = fs.step(0)
        assert np.array_equal(obs, ref.step(gray[i], fs_term[i])), i
        assert done == fs_term[i]
