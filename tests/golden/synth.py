"""Deterministic synthetic inputs shared by the golden-vector generator and the tests
(numpy default_rng with fixed seeds; identical bytes for oracle, CPU and GPU sides)."""
import numpy as np


def raw_frames(seed, count, kind="mixed"):
    """uint8 [count, 210, 160, 3] Atari-shaped raw frames.

    kind: "uniform" (iid bytes), "palette" (5-colour sparse, 90% background), "blocky"
    (10x10 flat blocks), "binary" (0/255), or "mixed" (cycles through all four).
    """
    rng = np.random.default_rng(seed)
    out = np.empty((count, 210, 160, 3), np.uint8)
    kinds = ["uniform", "palette", "blocky", "binary"]
    for i in range(count):
        k = kinds[i % 4] if kind == "mixed" else kind
        if k == "uniform":
            out[i] = rng.integers(0, 256, (210, 160, 3), dtype=np.uint8)
        elif k == "palette":
            pal = rng.integers(0, 256, (5, 3), dtype=np.uint8)
            pal[0] = 0
            idx = rng.choice(5, (210, 160), p=[.9, .025, .025, .025, .025])
            out[i] = pal[idx]
        elif k == "blocky":
            pal = rng.integers(0, 256, (5, 3), dtype=np.uint8)
            idx = np.kron(rng.integers(0, 5, (21, 16)), np.ones((10, 10), np.int64))
            out[i] = pal[idx]
        else:
            out[i] = (rng.integers(0, 2, (210, 160, 3)) * 255).astype(np.uint8)
    return out


def rollout(seed, num_envs, num_steps, num_actions=4, terminal_prob=0.05, obs_kind="uniform"):
    """The five train-step inputs of ActorCriticModel (model.py:97-105) with synthetic content."""
    rng = np.random.default_rng(seed)
    shape = (num_envs, num_steps, 84, 84, 4)
    if obs_kind == "uniform":
        obs = rng.integers(0, 256, shape, dtype=np.uint8)
        boot = rng.integers(0, 256, (num_envs, 84, 84, 4), dtype=np.uint8)
    else:  # sparse palette, Breakout-like
        pal = np.array([0, 52, 87, 142, 200], np.uint8)
        obs = pal[rng.choice(5, shape, p=[.9, .025, .025, .025, .025])]
        boot = pal[rng.choice(5, (num_envs, 84, 84, 4), p=[.9, .025, .025, .025, .025])]
    actions = rng.integers(0, num_actions, (num_envs, num_steps)).astype(np.uint8)
    rewards = rng.choice(np.array([-1, 0, 0, 0, 1], np.float32), (num_envs, num_steps))
    terminals = rng.random((num_envs, num_steps)) < terminal_prob
    return dict(observations=obs, bootstrap_observations=boot, actions=actions,
                rewards=rewards.astype(np.float32), terminals=terminals)


def fisher_samples(seed, num_rows, logits=None, num_actions=4):
    """Injected Fisher samples: y_hat uniform-ish ints and eps ~ N(0,1).

    When logits are given the labels are drawn from softmax(logits) (what kfac does); otherwise
    uniform labels (still a valid parity input: both sides consume the same labels)."""
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal(num_rows).astype(np.float32)
    if logits is None:
        y = rng.integers(0, num_actions, num_rows).astype(np.int32)
    else:
        z = np.asarray(logits, np.float64).reshape(num_rows, -1)
        p = np.exp(z - z.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        u = rng.random(num_rows)
        y = np.minimum((p.cumsum(1) < u[:, None]).sum(1), z.shape[1] - 1).astype(np.int32)
    return y, eps
