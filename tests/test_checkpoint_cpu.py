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
