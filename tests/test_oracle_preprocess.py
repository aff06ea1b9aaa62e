"""G1/G2: the preprocessing / frame-stack oracle against (a) the real cv2 the reference calls and
(b) golden vectors produced by the reference's own wrapper classes (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import synth
from oracle import preprocess as P


def test_gray_exhaustive_sample_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), P.rgb_to_gray(img))
    # every grey level of each pure channel
    ramp = np.zeros((3, 256, 3), np.uint8)
    for c in range(3):
        ramp[c, :, c] = np.arange(256)
    assert np.array_equal(cv2.cvtColor(ramp, cv2.COLOR_RGB2GRAY), P.rgb_to_gray(ramp))


@pytest.mark.parametrize("kind", ["uniform", "palette", "blocky", "binary"])
def test_area_resize_bit_exact_against_cv2(kind):
    cv2 = pytest.importorskip("cv2")
    frames = synth.raw_frames(7, 6, kind)
    for f in frames:
        g = cv2.cvtColor(f, cv2.COLOR_RGB2GRAY)
        ref = cv2.resize(g, (84, 84), interpolation=cv2.INTER_AREA)
        assert np.array_equal(ref, P.resize_area_84(g))


def test_tap_tables_shape():
    xs, xw, xn, ys, yw, yn = P.fixed_tap_tables()
    assert xn.sum() == 240 and yn.sum() == 252
    assert (xn == 3).sum() == 72 and (xn == 2).sum() == 12 and (yn == 3).all()
    np.testing.assert_allclose(xw[0, :2], [0.525, 0.475], rtol=1e-6)
    np.testing.assert_allclose(yw[0], [0.4, 0.4, 0.2], rtol=1e-6)


def test_preprocess_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    frames = synth.raw_frames(int(g["seed"]), int(g["count"]), "mixed")
    got = np.stack([P.preprocess_frame(f) for f in frames])
    assert got.dtype == np.uint8 and np.array_equal(got, g["observation"])
    assert np.array_equal(P.frame_max(frames[2], frames[3]), g["skip_max_23"])
    assert np.array_equal(P.frame_max(frames[4], frames[5]), g["skip_max_45"])
    assert np.array_equal(frames[0], g["skip_single_0"])


def _replay(g, e, stack_fn):
    """Re-drive env e of the framestack golden through `stack_fn` using the recorded raw-frame trace."""
    frames = synth.raw_frames(int(g["seeds"][e]), int(g["frames_per_env"]), "mixed")
    trace = g["trace_env%d" % e]
    return frames, trace


def test_framestack_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "framestack.npz"))
    for e in range(2):
        frames, trace = _replay(g, e, None)
        fs = P.FrameStack(4)
        step = 0
        outs = []
        first = None
        pending_terminal = False
        i = 0
        while i < len(trace):
            a, b, is_reset = trace[i]
            if is_reset:
                obs = fs.reset(P.preprocess_frame(frames[a]))
                if first is None:
                    first = obs.copy()
            else:
                frame = P.preprocess_frame(P.frame_max(frames[a], frames[b]))
                term = bool(g["terminals"][step, e])
                outs.append(fs.step(frame, term).copy())
                step += 1
            i += 1
        assert np.array_equal(first, g["reset_observation"][e])
        assert np.array_equal(np.stack(outs), g["observations"][:, e])
        del pending_terminal


def test_batched_stack_step_matches_per_env_class(golden_dir):
    g = np.load(os.path.join(golden_dir, "framestack.npz"))
    num_steps = g["observations"].shape[0]
    frames = [synth.raw_frames(int(g["seeds"][e]), int(g["frames_per_env"]), "mixed") for e in range(2)]
    traces = [list(map(tuple, g["trace_env%d" % e])) for e in range(2)]
    pos = [0, 0]
    stacks = np.zeros((2, 84, 84, 4), np.uint8)
    # initial reset
    for e in range(2):
        a, b, is_reset = traces[e][pos[e]]
        assert is_reset
        stacks[e] = np.repeat(P.preprocess_frame(frames[e][a]), 4, axis=-1)
        pos[e] += 1
    for t in range(num_steps):
        raw_a = np.zeros((2, 210, 160, 3), np.uint8)
        raw_b = np.zeros_like(raw_a)
        reset_raw = np.zeros_like(raw_a)
        reset_mask = np.zeros(2, bool)
        for e in range(2):
            a, b, is_reset = traces[e][pos[e]]
            if is_reset:
                reset_mask[e] = True
                reset_raw[e] = frames[e][a]
                pos[e] += 1
                a, b, is_reset = traces[e][pos[e]]
            raw_a[e], raw_b[e] = frames[e][a], frames[e][b]
            pos[e] += 1
        stacks = P.batched_stack_step(stacks, raw_a, raw_b, g["terminals"][t], reset_mask, reset_raw)
        assert np.array_equal(stacks, g["observations"][t])
