"""G3 on the GPU: K-RET against the reference golden and the oracle (tolerance 1e-5, north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import returns as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run(r, t, v, b, gamma=0.99):
    from actorcritic_b200 import ops
    tg, adv = ops.returns_adv(torch.from_numpy(r).cuda(), torch.from_numpy(t.astype(np.uint8)).cuda(),
                              torch.from_numpy(v).cuda(), torch.from_numpy(b).cuda(), gamma)
    return tg.cpu().numpy(), adv.cpu().numpy()


def _close(a, b):
    assert np.all(np.abs(a - b) <= TOL * np.maximum(1.0, np.abs(b)))


def test_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "returns.npz"))
    for ci in range(int(g["num_cases"])):
        r, t, b = g["rewards_%d" % ci], g["terminals_%d" % ci], g["bootstrap_%d" % ci]
        v = np.random.default_rng(ci).standard_normal(r.shape).astype(np.float32)
        tg, adv = _run(r, t, v, b, float(g["gamma"]))
        want = g["discounted_rewards_%d" % ci] + g["discounted_bootstrap_%d" % ci]
        _close(tg, want)
        _close(adv, want - v)


@pytest.mark.parametrize("e,t,p", [(32, 20, 0.05), (16, 5, 0.0), (256, 20, 1.0), (1, 1, 0.5), (4096, 20, 0.05), (3, 128, 0.1)])
def test_against_oracle_matrix_form(e, t, p):
    rng = np.random.default_rng(e + t)
    r = rng.standard_normal((e, t)).astype(np.float32)
    term = rng.random((e, t)) < p
    v = rng.standard_normal((e, t)).astype(np.float32)
    b = rng.standard_normal(e).astype(np.float32)
    tg, adv = _run(r, term, v, b)
    want = R.targets_matrix_form(r, term, b, 0.99, np.float64)
    _close(tg, want)
    _close(adv, want - v)
    # the kernel is the fp32 recursion with the same operation order: bit-exact against that form
    assert np.array_equal(tg, R.targets_recursive(r, term, b, 0.99, np.float32))
