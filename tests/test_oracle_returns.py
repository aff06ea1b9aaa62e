"""G3: returns / advantages / loss-value oracle against golden vectors produced by the reference's own
objectives._discount / _discount_bootstrap, plus form-equivalence and edge cases."""
import os

import numpy as np
import pytest

from oracle import returns as R

TOL = 1e-5   # north_star: <=1e-5 on returns/advantages


def _close(a, b):
    np.testing.assert_array_less(np.abs(a - b), TOL * np.maximum(1.0, np.abs(b)) + 1e-12)


def test_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "returns.npz"))
    gamma = float(g["gamma"])
    for ci in range(int(g["num_cases"])):
        r, t, b = g["rewards_%d" % ci], g["terminals_%d" % ci], g["bootstrap_%d" % ci]
        _close(R.discounted_rewards_matrix_form(r, t, gamma), g["discounted_rewards_%d" % ci])
        _close(R.bootstrap_factors(t, gamma) * b[:, None], g["discounted_bootstrap_%d" % ci])
        want = g["discounted_rewards_%d" % ci] + g["discounted_bootstrap_%d" % ci]
        _close(R.targets_matrix_form(r, t, b, gamma), want)
        _close(R.targets_recursive(r, t, b, gamma), want)
        _close(R.targets_recursive(r, t, b, gamma, np.float64), want)


def test_worked_example_survey_a3():
    term = np.zeros((1, 5), bool)
    term[0, 2] = True
    d = R.discount_matrix(term, 0.5)[0]
    # rewards @ D: column j collects rewards i>=j; rows are source steps
    want = np.array([[1, 0, 0, 0, 0], [.5, 1, 0, 0, 0], [.25, .5, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, .5, 1]], np.float32)
    np.testing.assert_array_equal(d, want)
    np.testing.assert_array_equal(R.bootstrap_factors(term, 0.5)[0], np.array([0, 0, 0, .25, .5], np.float32))


@pytest.mark.parametrize("e,t", [(32, 20), (16, 5), (256, 20), (1, 1)])
def test_recursion_equals_matrix_form_random(e, t):
    rng = np.random.default_rng(e * 100 + t)
    for p in (0.0, 0.05, 0.5, 1.0):
        r = rng.standard_normal((e, t)).astype(np.float32)
        term = rng.random((e, t)) < p
        b = rng.standard_normal(e).astype(np.float32)
        _close(R.targets_recursive(r, term, b, 0.99), R.targets_matrix_form(r, term, b, 0.99))


def test_losses_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "network.npz"))
    logits, values = g["logits"], g["values"]
    import synth
    batch = synth.rollout(int(g["seed"]), int(g["num_envs"]), int(g["num_steps"]), int(g["num_actions"]), terminal_prob=0.3)
    targets = R.targets_matrix_form(batch["rewards"], batch["terminals"], g["bootstrap_values"], 0.99, np.float64)
    out = R.a2c_losses(logits, values, batch["actions"], targets, 0.01)
    assert abs(out["policy_loss"] - float(g["policy_loss"])) < 1e-12
    assert abs(out["baseline_loss"] - float(g["baseline_loss"])) < 1e-12
    assert abs(out["mean_entropy"] - float(g["mean_entropy"])) < 1e-12
    np.testing.assert_allclose(out["log_prob"], g["log_prob"], atol=1e-12)
    np.testing.assert_allclose(out["entropy"], g["entropy"], atol=1e-12)
