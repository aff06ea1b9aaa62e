"""Checkpoint naming / index handling without a GPU (a2c_acktr.py:100-102,135-143,256-303)."""
import os

import numpy as np
import pytest

from actorcritic_b200 import checkpoint
from actorcritic_b200 import engine as eng


def test_filenames_follow_the_tensorflow_convention(tmp_path):
    assert checkpoint.checkpoint_filename("/x/model", 300) == "/x/model-300.npz"
    assert checkpoint.checkpoint_filename("/x/model") == "/x/model.npz"
    assert checkpoint.latest_checkpoint(str(tmp_path)) is None
    assert checkpoint.latest_checkpoint(str(tmp_path / "missing")) is None
    assert checkpoint.latest_checkpoint(None) is None


def test_latest_checkpoint_prefers_the_index_then_the_highest_step(tmp_path):
    for step in (100, 900, 1000):
        np.savez(str(tmp_path / ("model-%d.npz" % step)), x=np.zeros(1))
    assert checkpoint.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model-1000.npz")
    with open(tmp_path / "checkpoint", "w") as f:
        f.write('model_checkpoint_path: "model-900.npz"\n')
    assert checkpoint.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model-900.npz")
    os.remove(tmp_path / "model-900.npz")                      # a stale index falls back to the directory scan
    assert checkpoint.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model-1000.npz")


def test_variable_names_and_layouts_are_the_reference_ones():
    shapes = eng.param_shapes(4, 32)
    assert shapes["conv1/weights"] == (8, 8, 4, 32) and shapes["fc4/weights"] == (49 * 32, 512)      # HWIO / [in, out]
    params = eng.orthogonal_init(4, 32, seed=1)
    flat = eng.flatten_params(params, 4, 32)
    back = eng.unflatten_params(flat, 4, 32)
    assert all(np.array_equal(params[k], back[k]) for k in shapes)
    names = [k for k, _ in checkpoint._kfac_buffers()]
    assert "kfac/cov/A/heads" in names and "kfac/cov/G/fc_baseline" in names and "kfac/inv/A/fc_policy" in names
    assert len(names) == 5 + 3 * 6                            # 11 factors (5 input + 6 output) and 12 stored inverses


def test_saver_without_a_model_raises():
    checkpoint._last_model = None
    with pytest.raises(ValueError):
        checkpoint.Saver()


class _FakeEngine:
    """Host-memory stand-in with the Engine surface checkpoint.py uses (name mapping / flattening logic needs no device)."""

    def __init__(self, acktr, seed):
        import contextlib
        import torch
        self.config = eng.EngineConfig(num_envs=2, num_steps=2, conv3_filters=32, acktr=acktr)
        self.device = torch.device("cpu")
        self.num_params = int(sum(np.prod(s) for s in eng.param_shapes(4, 32).values()))
        g = torch.Generator().manual_seed(seed)
        self._flat = {k: torch.randn(self.num_params + 3, generator=g) for k in ("params", "velocity", "accum")}
        dims_a = {"conv1": 257, "conv2": 513, "conv3": 577, "fc4": 1569, "heads": 513}
        dims_l = {"conv1": (257, 32), "conv2": (513, 64), "conv3": (577, 32), "fc4": (1569, 512), "fc_policy": (513, 4),
                  "fc_baseline": (513, 1)}
        self._fac = {}
        for name, d in dims_a.items():
            self._fac[("sums", "A", name)] = torch.randn(8, 8, generator=g)      # small stand-ins: only names matter here
        for name, (da, dg) in dims_l.items():
            self._fac[("sums", "G", name)] = torch.randn(4, 4, generator=g)
            self._fac[("inv", "A", name)] = torch.randn(8, 8, generator=g)
            self._fac[("inv", "G", name)] = torch.randn(4, 4, generator=g)
        self.state = dict(global_step=123, num_cov_updates=45, inverses_valid=True)
        self.refreshed = 0
        self.on_stream = contextlib.nullcontext

    def buffer(self, name, dtype=None):
        return self._flat[name]

    def factor(self, kind, which, name):
        return self._fac[(kind, which, name)]

    def state_dict(self):
        sd = {k: v[:].clone() for k, v in self._flat.items()}
        sd.update(self.state)
        return sd

    def refresh_derived(self):
        self.refreshed += 1

    def set_state(self, gs, ncov=0, valid=False):
        self.state = dict(global_step=gs, num_cov_updates=ncov, inverses_valid=valid)


@pytest.mark.parametrize("acktr", [True, False])
def test_state_arrays_round_trip_through_the_reference_names(acktr):
    src, dst = _FakeEngine(acktr, 1), _FakeEngine(acktr, 2)
    arrays = checkpoint.state_to_arrays(src)
    n = src.num_params
    assert arrays["conv1/weights"].shape == (8, 8, 4, 32) and arrays["fc4/bias"].shape == (512,)
    if acktr:
        assert "fc4/weights/velocity" in arrays and "fc4/weights/Momentum" in arrays and "kfac/inv/G/fc_baseline" in arrays
        assert "fc4/weights/RMSProp" not in arrays
    else:
        assert "fc4/weights/RMSProp" in arrays and "fc4/weights/velocity" not in arrays and "kfac/cov/A/heads" not in arrays
    checkpoint.arrays_to_state(dst, arrays)
    assert dst.refreshed == 1 and dst.state == src.state
    for k in ("params", "accum") + (("velocity",) if acktr else ()):
        assert np.array_equal(dst._flat[k][:n].numpy(), src._flat[k][:n].numpy()), k
    if acktr:
        for key in src._fac:
            assert np.array_equal(dst._fac[key].numpy(), src._fac[key].numpy()), key
    # a checkpoint of the other optimizer only carries the variables over; slots and counters restart
    other = _FakeEngine(not acktr, 3)
    before = {k: v.clone() for k, v in other._flat.items()}
    checkpoint.arrays_to_state(other, arrays)
    assert np.array_equal(other._flat["params"][:n].numpy(), src._flat["params"][:n].numpy())
    assert np.array_equal(other._flat["accum"].numpy(), before["accum"].numpy())
    assert other.state == dict(global_step=123, num_cov_updates=0, inverses_valid=False)
    # architecture mismatch
    bad = dict(arrays)
    bad["meta/conv3_filters"] = np.int64(64)
    with pytest.raises(ValueError):
        checkpoint.arrays_to_state(dst, bad)
