"""Oracle for n-step discounted returns / advantages and the A2C loss values.

Follows objectives.py:123-154 (A2CObjective) and :178-214 (_discount, _discount_bootstrap).
Two forms are given for the returns: the reference's O(E*T^2) discount-matrix form (the thing
the reference literally evaluates) and the O(E*T) reverse recursion the CUDA kernel uses;
tests check they agree and that both agree with golden vectors made by running the
reference's own functions (tests/golden/make_golden.py).
"""
import numpy as np


def discount_matrix(terminals, discount_factor):
    """objectives.py:180-196: D[b,i,j] = gamma^(i-j) for i>=j with no terminal in [j, i-1], else 0.

    D is indexed [batch, source step i, target step j]; rewards @ D gives discounted rewards.
    float32, like the py_func output.
    """
    terminals = np.asarray(terminals, bool)
    e_count, t_count = terminals.shape
    gamma = np.float32(discount_factor)
    i = np.arange(t_count)[:, None]
    j = np.arange(t_count)[None, :]
    power = np.where(i >= j, i - j, 0).astype(np.float32)
    base = np.where(i >= j, gamma ** power, np.float32(0)).astype(np.float32)
    d = np.broadcast_to(base, (e_count, t_count, t_count)).copy()
    for b, t in np.argwhere(terminals):
        d[b, t + 1:, :t + 1] = 0.0
    return d


def discounted_rewards_matrix_form(rewards, terminals, discount_factor, dtype=np.float32):
    """objectives.py:198-202: squeeze(expand_dims(r,1) @ D, 1).  D is always the float32 py_func
    output; the matmul runs in `dtype` (float32 in the reference graph)."""
    rewards = np.asarray(rewards, dtype)
    d = discount_matrix(terminals, discount_factor).astype(dtype)
    return np.matmul(rewards[:, None, :], d)[:, 0, :].astype(dtype)


def bootstrap_factors(terminals, discount_factor):
    """objectives.py:209-211: f[b,t] = gamma^(T-t) * 1[no terminal in t..T-1]
    (int32 cumprod of ~terminals from the end, times gamma, float32 cumprod, flipped back)."""
    not_term = np.invert(np.asarray(terminals, bool))
    alive = np.cumprod(not_term[:, ::-1], axis=1, dtype=np.int32)
    f = np.cumprod(alive * discount_factor, axis=1, dtype=np.float32)
    return f[:, ::-1]


def targets_matrix_form(rewards, terminals, bootstrap_values, discount_factor, dtype=np.float32):
    """objectives.py:123-126."""
    return (discounted_rewards_matrix_form(rewards, terminals, discount_factor, dtype)
            + bootstrap_factors(terminals, discount_factor).astype(dtype)
            * np.asarray(bootstrap_values, dtype)[:, None])


def targets_recursive(rewards, terminals, bootstrap_values, discount_factor, dtype=np.float32):
    """R_T = V(s_T); R_t = r_t + gamma*(1-term_t)*R_{t+1}.  This is the form the CUDA
    kernel evaluates (one thread per environment, fp32, same operation order)."""
    rewards = np.asarray(rewards, dtype)
    term = np.asarray(terminals, bool)
    e_count, t_count = rewards.shape
    gamma = dtype(discount_factor)
    out = np.zeros((e_count, t_count), dtype)
    run = np.asarray(bootstrap_values, dtype).copy()
    for t in range(t_count - 1, -1, -1):
        run = np.where(term[:, t], dtype(0), run)
        run = (rewards[:, t] + gamma * run).astype(dtype)
        out[:, t] = run
    return out


def advantages(targets, values):
    """objectives.py:128-130."""
    return np.asarray(targets) - np.asarray(values)


def log_softmax(logits):
    z = np.asarray(logits)
    m = z.max(axis=-1, keepdims=True)
    s = z - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True))


def a2c_losses(logits, values, actions, targets, entropy_strength=0.01):
    """objectives.py:132-154 with policies.py:88-89,144.

    logits [E,T,A], values [E,T], actions [E,T] int, targets [E,T].
    Returns dict(policy_loss, baseline_loss, mean_entropy) (python floats of the array dtype).
    """
    logp_all = log_softmax(logits)
    p = np.exp(logp_all)
    act = np.asarray(actions).astype(np.int64)
    logp = np.take_along_axis(logp_all, act[..., None], axis=-1)[..., 0]
    entropy = -(p * logp_all).sum(axis=-1)
    adv = advantages(targets, values)
    mean_entropy = entropy.mean()
    policy_loss = -((adv * logp).mean() + entropy_strength * mean_entropy)
    baseline_loss = (np.square(np.asarray(targets) - np.asarray(values)) / 2.0).mean()
    return dict(policy_loss=policy_loss, baseline_loss=baseline_loss, mean_entropy=mean_entropy,
                log_prob=logp, entropy=entropy, advantage=adv)
