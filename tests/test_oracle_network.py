"""G4: the network/loss/gradient oracle against (a) golden vectors computed by autograd through the
REFERENCE'S OWN AtariModel + A2CObjective code (make_golden.py) and (b) torch autograd through the
oracle's own forward (independent check of the hand-derived backward)."""
import os

import numpy as np
import torch

import synth
from oracle import learner as L
from oracle import network as net


def _setup(g):
    batch = synth.rollout(int(g["seed"]), int(g["num_envs"]), int(g["num_steps"]), int(g["num_actions"]), terminal_prob=0.3)
    params = net.perturbed_params(int(g["num_actions"]), int(g["c3"]), seed=int(g["param_seed"]))
    return batch, params


def test_forward_and_losses_match_reference_code(golden_dir):
    g = np.load(os.path.join(golden_dir, "network.npz"))
    batch, params = _setup(g)
    # reference_cost=True evaluates the returns exactly as objectives.py:178-214 does (float32 matrices)
    ol = L.OracleLearner(params, int(g["num_actions"]), int(g["c3"]), acktr=False, reference_cost=True)
    info = ol.compute(batch, need_fisher=False)
    fast = L.OracleLearner(params, int(g["num_actions"]), int(g["c3"]), acktr=False).compute(batch, need_fisher=False)
    for k in ("policy_loss", "baseline_loss", "mean_entropy", "loss"):
        assert abs(float(fast["losses"][k]) - float(info["losses"][k])) < 1e-6, k   # fp32 discount-matrix rounding
    e, t = int(g["num_envs"]), int(g["num_steps"])
    np.testing.assert_allclose(info["fwd"]["logits"].numpy().reshape(e, t, -1), g["logits"], atol=1e-12)
    np.testing.assert_allclose(info["fwd"]["value"].numpy().reshape(e, t), g["values"], atol=1e-12)
    np.testing.assert_allclose(info["bootstrap_values"].numpy(), g["bootstrap_values"], atol=1e-12)
    for k in ("policy_loss", "baseline_loss", "mean_entropy"):
        assert abs(float(info["losses"][k]) - float(g[k])) < 1e-12, k
    assert abs(float(info["losses"]["loss"]) - float(g["shared_loss"])) < 1e-12


def test_gradients_match_reference_code(golden_dir):
    g = np.load(os.path.join(golden_dir, "network.npz"))
    batch, params = _setup(g)
    ol = L.OracleLearner(params, int(g["num_actions"]), int(g["c3"]), acktr=False, reference_cost=True)
    info = ol.compute(batch, need_fisher=False)
    for layer in net.LAYERS:
        dw, db = net.split_vmat(layer, info["grads"][layer], ol.params)
        key_w, key_b = "grad_%s_weights" % layer, "grad_%s_bias" % layer
        np.testing.assert_allclose(db.numpy(), g[key_b], atol=1e-13)
        if key_w in g.files:
            np.testing.assert_allclose(dw.numpy(), g[key_w], atol=1e-13)
        else:
            np.testing.assert_allclose(dw.numpy().reshape(-1)[::37], g[key_w + "_sample"], atol=1e-13)
            assert abs(float(dw.norm()) - float(g[key_w + "_norm"])) < 1e-12


def test_registrations_recorded_from_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "network.npz"))
    reg = [str(x) for x in g["registrations"]]
    assert len(reg) == 8 and bool(g["heads_share_inputs"])
    assert reg[0].startswith("conv2d inputs=(6, 84, 84, 4) outputs=(6, 20, 20, 32) strides=[1, 4, 4, 1] padding=VALID")
    assert reg[3].startswith("fully_connected inputs=(6, 1568) outputs=(6, 512)")
    assert reg[6] == "categorical logits=(2, 3, 4)" and reg[7] == "normal mean=(2, 3) var=1.0"


def test_hand_backward_equals_autograd():
    params = net.perturbed_params(4, 64, seed=1)
    batch = synth.rollout(3, 3, 4, terminal_prob=0.2, obs_kind="palette")
    pt = net.to_torch(params, torch.float64, requires_grad=True)
    n = 12
    fwd = net.forward(pt, batch["observations"].reshape(n, 84, 84, 4))
    boot = net.forward(pt, batch["bootstrap_observations"], build_policy=False)
    tg = net.targets_torch(batch["rewards"], batch["terminals"], boot["value"], float(np.float32(0.99))).reshape(n)
    net.a2c_loss(fwd["logits"], fwd["value"], batch["actions"].reshape(n), tg)["loss"].backward()
    ol = L.OracleLearner(params, 4, 64, acktr=False)
    info = ol.compute(batch, need_fisher=False)
    for layer in net.LAYERS:
        dw, db = net.split_vmat(layer, info["grads"][layer], ol.params)
        np.testing.assert_allclose(dw.numpy(), pt[layer + "/weights"].grad.numpy(), atol=1e-14)
        np.testing.assert_allclose(db.numpy(), pt[layer + "/bias"].grad.numpy(), atol=1e-14)


def test_fisher_output_grads_equal_autograd():
    rng = np.random.default_rng(0)
    logits = torch.tensor(rng.standard_normal((7, 4)), requires_grad=True)
    values = torch.tensor(rng.standard_normal(7), requires_grad=True)
    y, eps = synth.fisher_samples(0, 7, logits.detach().numpy())
    v_hat = (values + torch.tensor(eps, dtype=torch.float64)).detach()
    loss = -torch.log_softmax(logits, -1)[torch.arange(7), torch.tensor(y).long()].sum() \
        + (0.5 * (v_hat - values) ** 2).sum()
    loss.backward()
    dz, dv = net.fisher_output_grads(logits.detach(), values.detach(), y, eps)
    np.testing.assert_allclose(dz.numpy(), logits.grad.numpy(), atol=1e-14)
    np.testing.assert_allclose(dv.numpy(), values.grad.numpy(), atol=1e-6)   # eps is float32


def test_orthogonal_init():
    p = net.init_params(4, 32, 0)
    w = p["fc4/weights"].astype(np.float64)
    np.testing.assert_allclose(w.T @ w, 2.0 * np.eye(512), atol=1e-5)
    w = p["conv1/weights"].reshape(256, 32).astype(np.float64)
    np.testing.assert_allclose(w.T @ w, 2.0 * np.eye(32), atol=1e-5)
    w = p["fc_policy/weights"].astype(np.float64)
    np.testing.assert_allclose(w.T @ w, 1e-4 * np.eye(4), atol=1e-9)
    assert all(not p[k].any() for k in p if k.endswith("bias"))
