"""`Session` - what replaces `tf.Session` at the drop-in boundary: an opaque handle of a CUDA device
and stream that evaluates "fetches" against a feed_dict, exactly how the reference drives its learner:

    session.run([summary_op, global_step, optimize_op], feed_dict={model.observations_placeholder: ..., ...})
                                                                     (actorcritic/examples/atari/a2c_acktr.py:117-126)
    session.run(policy.sample, feed_dict={observations_placeholder: [E,1,84,84,4]})            (model.py:149-151)

There is no graph: fetches are small token objects (`Fetch`) created by the model / policy / objective; the session
groups them, copies the fed host arrays into the engine's device arena and launches the native kernels.  All
arithmetic is in libacx.so - a session cannot be created without a CUDA device.
"""
import numpy as np
import torch

from . import _lib


class Placeholder:
    """Named input slot (model.py:97-105).  `shape` uses None for the [environment, step] batch dimensions."""

    def __init__(self, name, dtype, shape):
        self.name, self.dtype, self.shape = name, np.dtype(dtype), tuple(shape)

    def __repr__(self):
        return "<Placeholder %s %s %s>" % (self.name, self.dtype, self.shape)


class Fetch:
    """Something `Session.run` can evaluate.  kind: one of
    policy_loss baseline_loss mean_entropy loss clip_coeff learning_rate global_step optimize
    sample mode logits log_prob entropy value bootstrap_values summary no_op"""

    def __init__(self, kind, owner, name=None):
        self.kind, self.owner, self.name = kind, owner, name or kind

    def __repr__(self):
        return "<Fetch %s>" % self.name


_GLOBAL_STEPS = []    # weak references to every GlobalStep (checkpoint restore before the learner engine exists)


def restore_global_steps(value):
    """checkpoint.Saver.restore: global steps that are not bound to an engine yet report the restored value (the
    reference reads `session.run(global_step)` right after `saver.restore`, a2c_acktr.py:100-104)."""
    for ref in list(_GLOBAL_STEPS):
        gs = ref()
        if gs is None:
            _GLOBAL_STEPS.remove(ref)
        elif gs._engine is None or not getattr(gs._engine, "is_learner", True):
            gs._pending = int(value)


class GlobalStep(Fetch):
    """`tf.train.get_or_create_global_step()` stand-in: fetchable; the value lives in the engine."""

    def __init__(self):
        import weakref
        super().__init__("global_step", None)
        self._engine = None
        self._pending = 0
        _GLOBAL_STEPS.append(weakref.ref(self))

    def bind(self, engine):
        self._engine = engine
        if self._pending and engine.global_step != self._pending:     # keep the engine's other schedule counters
            st = engine.get_state()
            engine.set_state(self._pending, st["num_cov_updates"], st["inverses_valid"])

    def eval(self):
        return self._engine.global_step if self._engine is not None else self._pending


_SCALARS = {"policy_loss": "policy_loss", "baseline_loss": "baseline_loss", "mean_entropy": "mean_entropy", "loss": "loss",
            "clip_coeff": "clip_coeff", "learning_rate": "learning_rate"}


class Session:
    def __init__(self, device=None, group=None):
        if not torch.cuda.is_available():
            raise _lib.AcxError("actorcritic_b200.Session needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.group = group          # torch.distributed process group for data-parallel learners (None = default)
        self.fisher_injection = None  # (labels int32 [N], eps f32 [N]) device tensors for parity runs

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    # -------------------------------------------------------------------------------------------
    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        for f in flist:
            if not isinstance(f, Fetch):
                raise TypeError("cannot fetch %r: not a value of this framework" % (f,))
        feed = {} if feed_dict is None else feed_dict
        for k in feed:
            if not isinstance(k, Placeholder):
                raise TypeError("feed_dict keys must be placeholders, got %r" % (k,))
        out = self._evaluate(flist, feed)
        return out[0] if single else out

    def _evaluate(self, flist, feed):
        # summary ops (actorcritic_b200.summary) are evaluated from the same step as the other fetches: their sources are
        # added to the work list, the serialized Summary is assembled at the end; no_op fetches yield None
        requested = flist
        summaries = [f for f in requested if f.kind == "summary"]
        flist = [f for f in requested if f.kind not in ("summary", "no_op")]
        for sm in summaries:
            for src in sm.sources():
                if all(src is not g for g in flist):
                    flist.append(src)
        results = self._evaluate_plain(flist, feed) if flist else {}
        for sm in summaries:
            results[id(sm)] = sm.build(results, feed)
        for f in requested:
            if f.kind == "no_op":
                results[id(f)] = None
        return [results[id(f)] for f in requested]

    def _evaluate_plain(self, flist, feed):
        kinds = {f.kind for f in flist}
        results = {}
        train_kinds = {"optimize", "optimize_separate", "policy_loss", "baseline_loss", "mean_entropy", "loss", "clip_coeff",
                       "learning_rate"}
        act_kinds = {"sample", "mode", "logits", "value"}
        if kinds & train_kinds:
            objective = next(f.owner for f in flist if f.kind in train_kinds and f.owner is not None)
            model = objective.model
            engine = model._engine_for_feed(self, feed, objective)
            batch = model._train_feed(feed)
            engine.load_batch(*batch)
            injected = self.fisher_injection or (None, None)
            if "optimize_separate" in kinds:       # objectives.py:31-54: two optimizers, two backward passes
                sep = objective._separate
                inc = (int(sep["policy_global_step"] is not None), int(sep["baseline_global_step"] is not None))
                scal = engine.update_separate(None, sep["policy"], sep["baseline"], inc)
            else:
                engine.phase1(injected[0], injected[1])
            if "optimize_separate" in kinds:
                pass
            elif "optimize" in kinds:
                engine.allreduce(self.group)
                engine.phase2()
                scal = engine.fetch_scalars()
            else:
                torch.cuda.current_stream(self.device).synchronize()
                b = engine.bucket[-4:].cpu().tolist()
                scal = dict(policy_loss=b[0], baseline_loss=b[1], mean_entropy=b[2], loss=b[3], clip_coeff=float("nan"),
                            learning_rate=float("nan"))
            for f in flist:
                if f.kind in _SCALARS:
                    results[id(f)] = np.float32(scal[_SCALARS[f.kind]])
                elif f.kind in ("optimize", "optimize_separate"):
                    results[id(f)] = None
            if "bootstrap_values" in kinds:
                n = engine.rows
                bv = engine.values[n:].cpu().numpy().copy()
                for f in flist:
                    if f.kind == "bootstrap_values":
                        results[id(f)] = bv
        if kinds & act_kinds:
            owner = next(f.owner for f in flist if f.kind in act_kinds)
            model = owner if hasattr(owner, "_act") else owner.model
            act = model._act(self, feed, greedy="mode" in kinds and "sample" not in kinds)
            for f in flist:
                if f.kind in act_kinds:
                    results[id(f)] = act["mode" if f.kind == "mode" else f.kind]
        for f in flist:
            if f.kind == "global_step":
                results[id(f)] = np.int64(f.eval())
            elif id(f) not in results:
                if f.kind == "bootstrap_values":
                    model = f.owner
                    results[id(f)] = model._bootstrap_only(self, feed)
                else:
                    raise NotImplementedError("fetch %r is not available outside the learner hot path" % f)
        return results
