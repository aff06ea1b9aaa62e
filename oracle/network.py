"""Oracle for the Nature-CNN forward/backward and the A2C objective (torch-CPU, dtype selectable).

Follows envs/atari/model.py:92-127 (normalise, flatten, two towers, head reshapes), :129-170
(parameter shapes, orthogonal init gains), :173-217 (layers), nn.py:37-52,88-110,114-126,
policies.py:86-89,144, objectives.py:123-154,78.

Everything is written as explicit im2col + matmul (patch order (kh,kw,cin)) so that the
same matrices the K-FAC factors need (patches P_l, pre-activation gradients g_l) fall out, and the
hand-derived backward can be checked against torch autograd (tests/test_oracle_network.py).
"""
import math

import numpy as np
import torch

LAYERS = ("conv1", "conv2", "conv3", "fc4", "fc_policy", "fc_baseline")
CONV_GEOM = {  # name: (kernel, stride, cin, in_hw, out_hw)
    "conv1": (8, 4, 4, 84, 20),
    "conv2": (4, 2, 32, 20, 9),
    "conv3": (3, 1, 64, 9, 7),
}


def conv_cout(name, c3):
    return {"conv1": 32, "conv2": 64, "conv3": c3}[name]


def param_shapes(num_actions=4, c3=32):
    """envs/atari/model.py:137-170 ; nn.py:31-32,81-83."""
    return {
        "conv1/weights": (8, 8, 4, 32), "conv1/bias": (32,),
        "conv2/weights": (4, 4, 32, 64), "conv2/bias": (64,),
        "conv3/weights": (3, 3, 64, c3), "conv3/bias": (c3,),
        "fc4/weights": (49 * c3, 512), "fc4/bias": (512,),
        "fc_policy/weights": (512, num_actions), "fc_policy/bias": (num_actions,),
        "fc_baseline/weights": (512, 1), "fc_baseline/bias": (1,),
    }


def orthogonal(shape, gain, rng):
    """tf.orthogonal_initializer semantics: QR of a normal [prod(shape[:-1]), shape[-1]]
    matrix (transposed if rows < cols), sign-fixed by diag(R), scaled by gain."""
    rows = int(np.prod(shape[:-1]))
    cols = int(shape[-1])
    flat = (max(rows, cols), min(rows, cols))
    a = rng.standard_normal(flat)
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q).reshape(shape).astype(np.float32)


def init_params(num_actions=4, c3=32, seed=0):
    """envs/atari/model.py:132-135: gains sqrt(2) trunk, 0.01 policy, 1.0 value; zero biases."""
    rng = np.random.default_rng(seed)
    gains = {"conv1": math.sqrt(2.0), "conv2": math.sqrt(2.0), "conv3": math.sqrt(2.0), "fc4": math.sqrt(2.0),
             "fc_policy": 0.01, "fc_baseline": 1.0}
    params = {}
    for key, shape in param_shapes(num_actions, c3).items():
        layer, kind = key.split("/")
        params[key] = orthogonal(shape, gains[layer], rng) if kind == "weights" else np.zeros(shape, np.float32)
    return params


def perturbed_params(num_actions=4, c3=32, seed=0, bias_scale=0.05, policy_gain=0.5):
    """Parity-test weights: orthogonal init but with non-zero biases and a larger policy gain so that
    every term of the gradient (bias gradients, softmax away from uniform) is exercised."""
    rng = np.random.default_rng(seed + 1000)
    p = init_params(num_actions, c3, seed)
    for key in p:
        if key.endswith("/bias"):
            p[key] = (bias_scale * rng.standard_normal(p[key].shape)).astype(np.float32)
    p["fc_policy/weights"] = (p["fc_policy/weights"] * (policy_gain / 0.01)).astype(np.float32)
    return p


def to_torch(params, dtype=torch.float64, requires_grad=False):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in params.items()}


def im2col(x, k, s):
    """x [N,H,W,C] -> patches [N*oh*ow, k*k*C] ordered (kh, kw, cin)."""
    n, h, w, c = x.shape
    p = x.unfold(1, k, s).unfold(2, k, s)            # [N, oh, ow, C, kh, kw]
    p = p.permute(0, 1, 2, 4, 5, 3)                   # [N, oh, ow, kh, kw, C]
    oh, ow = p.shape[1], p.shape[2]
    return p.reshape(n * oh * ow, k * k * c), oh, ow


def col2im(dp, n, h, w, c, k, s):
    """Adjoint of im2col: dp [N*oh*ow, k*k*C] -> dx [N,H,W,C]."""
    oh = (h - k) // s + 1
    ow = (w - k) // s + 1
    dp = dp.reshape(n, oh, ow, k, k, c)
    dx = torch.zeros((n, h, w, c), dtype=dp.dtype)
    for ky in range(k):
        for kx in range(k):
            dx[:, ky:ky + s * oh:s, kx:kx + s * ow:s, :] += dp[:, :, :, ky, kx, :]
    return dx


def forward(params, obs_u8, build_policy=True):
    """AtariModel._build_layers (:173-217) on flat rows.  obs_u8: [R,84,84,4] uint8 (numpy or torch).

    Returns dict with inputs/patches/pre-activations/activations per layer.  dtype = dtype of params.
    """
    dtype = params["conv1/weights"].dtype
    x = torch.as_tensor(np.asarray(obs_u8)).to(dtype) / 255.0       # :93
    r = x.shape[0]
    out = {"rows": r}
    a = x
    for name in ("conv1", "conv2", "conv3"):
        k, s, cin, _, _ = CONV_GEOM[name]
        w = params[name + "/weights"]
        p, oh, ow = im2col(a, k, s)
        pre = p @ w.reshape(k * k * cin, -1) + params[name + "/bias"]      # nn.py:108-110
        act = torch.relu(pre)
        out[name] = dict(patches=p, pre=pre, act=act, oh=oh, ow=ow)
        a = act.reshape(r, oh, ow, -1)
    flat = a.reshape(r, -1)                                                  # nn.py:125-126 (h,w,c)
    pre4 = flat @ params["fc4/weights"] + params["fc4/bias"]
    act4 = torch.relu(pre4)
    out["fc4"] = dict(inputs=flat, pre=pre4, act=act4)
    if build_policy:
        out["logits"] = act4 @ params["fc_policy/weights"] + params["fc_policy/bias"]
    out["value"] = (act4 @ params["fc_baseline/weights"] + params["fc_baseline/bias"])[:, 0]
    out["heads_inputs"] = act4
    return out


def targets_torch(rewards, terminals, bootstrap_values, gamma):
    """Reverse recursion in the dtype of bootstrap_values."""
    dtype = bootstrap_values.dtype
    rewards = torch.as_tensor(np.asarray(rewards)).to(dtype)
    term = torch.as_tensor(np.asarray(terminals, bool))
    t_count = rewards.shape[1]
    run = bootstrap_values.detach().clone()
    cols = [None] * t_count
    for t in range(t_count - 1, -1, -1):
        run = torch.where(term[:, t], torch.zeros_like(run), run)
        run = rewards[:, t] + gamma * run
        cols[t] = run
    return torch.stack(cols, dim=1)


def a2c_loss(logits, values, actions, targets, beta=0.01, value_weight=0.5):
    """objectives.py:128-154,78.  logits [N,A], values [N], actions [N], targets [N] (flat rows)."""
    logp_all = torch.log_softmax(logits, dim=-1)
    p = torch.exp(logp_all)
    act = torch.as_tensor(np.asarray(actions)).long()
    logp = torch.gather(logp_all, 1, act[:, None])[:, 0]
    entropy = -(p * logp_all).sum(-1)
    adv = (targets - values).detach()
    mean_entropy = entropy.mean()
    policy_loss = -((adv * logp).mean() + beta * mean_entropy)
    baseline_loss = (((targets.detach() - values) ** 2) / 2.0).mean()
    loss = policy_loss + value_weight * baseline_loss
    return dict(loss=loss, policy_loss=policy_loss, baseline_loss=baseline_loss, mean_entropy=mean_entropy,
                advantage=adv, log_prob=logp, entropy=entropy)


def output_grads(logits, values, actions, targets, beta=0.01, value_weight=0.5):
    """Closed-form dL/dlogits [N,A], dL/dvalue [N]."""
    n = logits.shape[0]
    logp_all = torch.log_softmax(logits, dim=-1)
    p = torch.exp(logp_all)
    ent = -(p * logp_all).sum(-1, keepdim=True)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(np.asarray(actions)).long(), logits.shape[1]).to(logits.dtype)
    adv = (targets - values).detach()[:, None]
    dz = -(adv / n) * (onehot - p) + (beta / n) * p * (logp_all + ent)
    dv = -value_weight * (targets - values) / n
    return dz, dv


def fisher_output_grads(logits, values, y_hat, eps):
    """L_s = -sum log p(y_hat) - sum log N(v_hat; V, 1), v_hat = V + eps.
    d/dz = p - onehot(y_hat);  d/dV = V - v_hat = -eps."""
    p = torch.softmax(logits, dim=-1)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(np.asarray(y_hat)).long(), logits.shape[1]).to(logits.dtype)
    return p - onehot, -torch.as_tensor(np.asarray(eps)).to(values.dtype)


def backward(params, fwd, dlogits, dvalue, masks=None):
    """Hand-derived backward through the trunk for output gradients (dlogits [R,A], dvalue [R]).

    Returns (grads, pre_grads): grads[name] = V_l = [dW reshaped [K_l, C_l]; db] as one [K_l+1, C_l]
    matrix (SURVEY A.5 'Precondition'), pre_grads[name] = dL/d(pre-activation) [rows_l, C_l].

    masks: optional {layer: bool [rows_l, C_l]} replacing the ReLU derivative `pre > 0` - the parity tests pass the
    masks of the implementation under test, so that a unit whose pre-activation is within rounding of zero (and takes
    the other branch there) does not drown the arithmetic comparison (tests/test_gpu_learner.py).
    """
    def relu_mask(name, pre):
        return (pre > 0) if masks is None or name not in masks else torch.as_tensor(masks[name]).reshape(pre.shape)

    grads, g = {}, {}
    act4 = fwd["heads_inputs"]
    g["fc_policy"] = dlogits
    g["fc_baseline"] = dvalue[:, None]
    grads["fc_policy"] = torch.cat([act4.T @ dlogits, dlogits.sum(0, keepdim=True)], 0)
    grads["fc_baseline"] = torch.cat([act4.T @ g["fc_baseline"], g["fc_baseline"].sum(0, keepdim=True)], 0)
    d_act4 = dlogits @ params["fc_policy/weights"].T + g["fc_baseline"] @ params["fc_baseline/weights"].T
    g4 = d_act4 * relu_mask("fc4", fwd["fc4"]["pre"])
    g["fc4"] = g4
    grads["fc4"] = torch.cat([fwd["fc4"]["inputs"].T @ g4, g4.sum(0, keepdim=True)], 0)
    d_flat = g4 @ params["fc4/weights"].T                       # [R, 49*c3]
    r = fwd["rows"]
    d_act = d_flat.reshape(r * 49, -1)
    for name in ("conv3", "conv2", "conv1"):
        k, s, cin, in_hw, out_hw = CONV_GEOM[name]
        layer = fwd[name]
        gl = d_act * relu_mask(name, layer["pre"])
        g[name] = gl
        grads[name] = torch.cat([layer["patches"].T @ gl, gl.sum(0, keepdim=True)], 0)
        if name != "conv1":
            w2d = params[name + "/weights"].reshape(k * k * cin, -1)
            dp = gl @ w2d.T
            dx = col2im(dp, r, in_hw, in_hw, cin, k, s)
            d_act = dx.reshape(r * in_hw * in_hw, cin)
    return grads, g


def split_vmat(name, vmat, params):
    """[K+1, C] -> (dW in the variable's shape, db)."""
    w = params[name + "/weights"]
    return vmat[:-1].reshape(w.shape), vmat[-1].reshape(params[name + "/bias"].shape)


def join_vmat(name, params):
    w = params[name + "/weights"]
    return torch.cat([w.reshape(-1, w.shape[-1]), params[name + "/bias"].reshape(1, -1)], 0)
