"""Checkpoint / resume of the learner state under the reference's variable names.

The reference saves with `tf.train.Saver()` over all global variables and restores the latest checkpoint of a
directory (actorcritic/examples/atari/a2c_acktr.py:100-102 `load_model`, :135-143 `save_model` every 100th step and on
Ctrl-C, :256-303).  This module keeps that call shape:

    saver = checkpoint.Saver(model)                       # tf.train.Saver()
    path = checkpoint.latest_checkpoint(directory)        # tf.train.latest_checkpoint
    if path is not None: saver.restore(session, path)     # saver.restore(session, path)
    saver.save(session, directory + '/' + model_name, step)   # saver.save(session, prefix, global_step)

File format: one `.npz` per checkpoint, `<prefix>-<step>.npz`, plus a text index file `checkpoint` in the same
directory naming the latest one (the TensorFlow convention).  Array names and layouts follow the reference's graph
(envs/atari/model.py:137-170, nn.py:31-32,81-83) so that a converter from / to a real TensorFlow checkpoint is a
rename: `conv1/weights` is HWIO, `fc4/weights` is [in, out]; optimizer slots carry the slot name TensorFlow gives them
(`<var>/Momentum`, `<var>/RMSProp`, K-FAC `<var>/velocity`); K-FAC state is stored per factor (`kfac/cov/A/<layer>`,
`kfac/cov/G/<layer>`, `kfac/inv/...`) together with the schedule counters.
"""
import os
import re
import weakref

import numpy as np
import torch

from . import engine as eng

FORMAT_VERSION = 1
_INDEX = "checkpoint"
_last_model = None


def _register_model(model):
    """Called by AtariModel.__init__: `Saver()` without arguments saves the most recently built model, like
    `tf.train.Saver()` saves the variables of the default graph."""
    global _last_model
    _last_model = weakref.ref(model)


def _kfac_buffers():
    """(checkpoint name, engine.factor arguments) of every K-FAC matrix: running covariance sums per factor (the two
    heads share one input factor, envs/atari/model.py:243,246) and the stored inverses per layer."""
    out = []
    for name in eng.A_FACTORS:
        out.append(("kfac/cov/A/" + name, ("sums", "A", name)))
    for name in eng.LAYERS:
        out.append(("kfac/cov/G/" + name, ("sums", "G", name)))
        out.append(("kfac/inv/A/" + name, ("inv", "A", name)))
        out.append(("kfac/inv/G/" + name, ("inv", "G", name)))
    return out


def state_to_arrays(engine):
    """Engine state -> {name: numpy array} in the reference's naming."""
    cfg = engine.config
    sd = engine.state_dict()
    out = {}
    params = eng.unflatten_params(sd["params"].numpy(), cfg.num_actions, cfg.conv3_filters)
    out.update(params)
    velocity = eng.unflatten_params(sd["velocity"].numpy(), cfg.num_actions, cfg.conv3_filters)
    accum = eng.unflatten_params(sd["accum"].numpy(), cfg.num_actions, cfg.conv3_filters)
    for k in params:
        if cfg.acktr:
            out[k + "/velocity"] = velocity[k]          # kfac's momentum slot
            out[k + "/Momentum"] = accum[k]              # cold optimizer: tf.train.MomentumOptimizer slot
        else:
            out[k + "/RMSProp"] = accum[k]               # tf.train.RMSPropOptimizer `ms` slot (initialised to ones)
    if cfg.acktr:
        for key, args in _kfac_buffers():
            out[key] = engine.factor(*args).cpu().numpy().copy()
    out["global_step"] = np.int64(sd["global_step"])
    out["kfac/num_cov_updates"] = np.int64(sd["num_cov_updates"])
    out["kfac/inverses_valid"] = np.bool_(sd["inverses_valid"])
    out["meta/format_version"] = np.int64(FORMAT_VERSION)
    out["meta/acktr"] = np.bool_(cfg.acktr)
    out["meta/num_actions"] = np.int64(cfg.num_actions)
    out["meta/conv3_filters"] = np.int64(cfg.conv3_filters)
    return out


def arrays_to_state(engine, arrays, params_only=False):
    """Inverse of `state_to_arrays`: writes the arrays into the engine (device) and re-derives the operand planes.
    params_only: an acting-only engine (built by sample_actions before the first train step) takes the variables and the
    step counter; the optimizer state waits in `model._pending_checkpoint` for the learner engine."""
    cfg = engine.config
    if hasattr(engine, "wait_pending_ema"):
        engine.wait_pending_ema()   # a split exchange's EMA may still be writing the running sums
    if int(arrays["meta/num_actions"]) != cfg.num_actions or int(arrays["meta/conv3_filters"]) != cfg.conv3_filters:
        raise ValueError("checkpoint is for %d actions / conv3=%d, the model has %d / %d" % (
            int(arrays["meta/num_actions"]), int(arrays["meta/conv3_filters"]), cfg.num_actions, cfg.conv3_filters))
    shapes = eng.param_shapes(cfg.num_actions, cfg.conv3_filters)
    for k, shape in shapes.items():
        if k not in arrays:
            raise KeyError("checkpoint has no variable %r" % k)
        if tuple(arrays[k].shape) != shape:
            raise ValueError("variable %s: checkpoint shape %s, model shape %s" % (k, arrays[k].shape, shape))
    flat = lambda suffix: eng.flatten_params({k: arrays[k + suffix] for k in shapes}, cfg.num_actions, cfg.conv3_filters)
    dev = engine.device
    with engine.on_stream():
        engine.buffer("params", torch.float32)[:engine.num_params].copy_(torch.from_numpy(flat("")).to(dev))
        same_kind = (bool(arrays["meta/acktr"]) == bool(cfg.acktr) and not bool(arrays.get("meta/params_only", False))
                     and not params_only)
        if same_kind:                       # slots only carry over between runs of the same optimizer
            if cfg.acktr:
                engine.buffer("velocity", torch.float32)[:engine.num_params].copy_(
                    torch.from_numpy(flat("/velocity")).to(dev))
                engine.buffer("accum", torch.float32)[:engine.num_params].copy_(
                    torch.from_numpy(flat("/Momentum")).to(dev))
                for key, args in _kfac_buffers():
                    engine.factor(*args).copy_(torch.from_numpy(arrays[key]).to(dev))
            else:
                engine.buffer("accum", torch.float32)[:engine.num_params].copy_(
                    torch.from_numpy(flat("/RMSProp")).to(dev))
    engine.refresh_derived()
    if same_kind:
        engine.set_state(int(arrays["global_step"]), int(arrays["kfac/num_cov_updates"]),
                         bool(arrays["kfac/inverses_valid"]))
    else:
        engine.set_state(int(arrays["global_step"]), 0, False)


def checkpoint_filename(save_path, global_step=None):
    return "%s-%d.npz" % (save_path, int(global_step)) if global_step is not None else save_path + ".npz"


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the path recorded in `<dir>/checkpoint`, else the highest-step `*.npz`, else None."""
    if checkpoint_dir is None or not os.path.isdir(checkpoint_dir):
        return None
    index = os.path.join(checkpoint_dir, _INDEX)
    if os.path.exists(index):
        with open(index) as f:
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', f.read())
        if m:
            path = m.group(1)
            if not os.path.isabs(path):
                path = os.path.join(checkpoint_dir, path)
            if os.path.exists(path):
                return path
    best, best_step = None, -1
    for name in os.listdir(checkpoint_dir):
        m = re.match(r"(.+)-(\d+)\.npz$", name)
        if m and int(m.group(2)) > best_step:
            best, best_step = os.path.join(checkpoint_dir, name), int(m.group(2))
    return best


class Saver:
    """`tf.train.Saver` for this framework's learner.  `max_to_keep` older checkpoints are deleted like TensorFlow does
    (default 5)."""

    def __init__(self, model=None, max_to_keep=5):
        if model is None:
            model = _last_model() if _last_model is not None else None
            if model is None:
                raise ValueError("Saver(): no model has been built yet")
        self._model = model
        self._max_to_keep = max_to_keep
        self._kept = []

    def save(self, session, save_path, global_step=None):
        if isinstance(global_step, torch.Tensor):
            global_step = int(global_step.item())
        elif hasattr(global_step, "eval") and not isinstance(global_step, (int, np.integer)):
            global_step = int(global_step.eval())
        path = checkpoint_filename(save_path, global_step)
        directory = os.path.dirname(os.path.abspath(path))
        os.makedirs(directory, exist_ok=True)
        engine = self._model.engine
        pending = getattr(self._model, "_pending_checkpoint", None)
        if engine is not None and not getattr(engine, "is_learner", True) and pending is not None:
            # only an acting engine exists so far (restore -> interact -> save, e.g. the reference's Ctrl-C handler during the
            # first rollout): the restored optimizer state is still waiting for the learner engine - save THAT, with the
            # engine's current variables
            arrays = dict(pending)
            arrays.update(engine.get_params())
        elif engine is not None:
            arrays = state_to_arrays(engine)
        else:
            # no train step has run yet (the learner is built from the first feed): the variables are all there is
            arrays = dict(pending) if pending is not None else dict(self._model.get_variables())
            if pending is None:
                arrays.update({"global_step": np.int64(0), "kfac/num_cov_updates": np.int64(0),
                               "kfac/inverses_valid": np.bool_(False), "meta/format_version": np.int64(FORMAT_VERSION),
                               "meta/acktr": np.bool_(False), "meta/params_only": np.bool_(True),
                               "meta/num_actions": np.int64(self._model.num_actions),
                               "meta/conv3_filters": np.int64(self._model.conv3_num_filters)})
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, **arrays)
        os.replace(tmp, path)                                   # a crash never leaves a half-written checkpoint
        self._kept.append(path)
        while self._max_to_keep and len(self._kept) > self._max_to_keep:
            old = self._kept.pop(0)
            if old != path and os.path.exists(old):
                os.remove(old)
        with open(os.path.join(directory, _INDEX), "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
            for k in self._kept:
                f.write('all_model_checkpoint_paths: "%s"\n' % os.path.basename(k))
        return path

    def restore(self, session, save_path):
        if not os.path.exists(save_path):
            raise FileNotFoundError(save_path)
        with np.load(save_path) as z:
            arrays = {k: z[k] for k in z.files}
        if int(arrays.get("meta/format_version", -1)) != FORMAT_VERSION:
            raise ValueError("%s: unknown checkpoint format" % save_path)
        from . import session as _session
        _session.restore_global_steps(int(arrays["global_step"]))      # `session.run(global_step)` right after a restore
        engine = self._model.engine
        if engine is None or not getattr(engine, "is_learner", True):
            # the learner is built lazily from the first train feed: keep the arrays and apply them then; the variables
            # (and the step counter) go to the model / the acting-only engine now
            self._model._pending_checkpoint = arrays
            self._model.set_variables({k: arrays[k] for k in eng.param_shapes(self._model.num_actions,
                                                                               self._model.conv3_num_filters)})
            if engine is not None:
                arrays_to_state(engine, arrays, params_only=True)
            return
        arrays_to_state(engine, arrays)
