"""actorcritic/envs/atari/model.py: AtariModel - the Nature-CNN actor-critic for 84x84x4 uint8 observations.

    conv1 8x8/4 4->32, conv2 4x4/2 32->64, conv3 3x3/1 64->conv3_num_filters, fc4 ->512, fc_policy ->A, fc_baseline ->1
    (envs/atari/model.py:129-217; VALID padding, ReLU, NHWC, HWIO kernels, orthogonal init gains sqrt2 / 0.01 / 1.0)

Variables keep the reference's names ({conv1,conv2,conv3,fc4,fc_policy,fc_baseline}/{weights,bias}) and layouts, so
weights can be exchanged with a reference checkpoint.  Forward, backward and the K-FAC update run in libacx's
learner engine, which is created on the first train step, when the [environments, steps] batch shape is known.
"""
import numpy as np
import torch

from ... import checkpoint
from ... import engine as eng
from ... import spaces
from ...baselines import StateValueFunction
from ...model import ActorCriticModel
from ...policies import SoftmaxPolicy
from ...session import Fetch


class AtariModel(ActorCriticModel):
    def __init__(self, observation_space, action_space, conv3_num_filters=64, random_seed=None, name=None):
        """envs/atari/model.py:45.  conv3_num_filters: 64 by default, 32 for ACKTR (a2c_acktr.py:52)."""
        super().__init__(observation_space, action_space)
        if tuple(observation_space.shape) != eng.OBS_SHAPE:
            raise ValueError("AtariModel expects 84x84x4 observations, got %s" % (tuple(observation_space.shape),))
        if not spaces.is_discrete(action_space):
            raise TypeError("Unsupported space")
        self.num_actions = int(action_space.n)
        self.conv3_num_filters = int(conv3_num_filters)
        self.random_seed = random_seed
        self._params = eng.orthogonal_init(self.num_actions, self.conv3_num_filters, random_seed)
        self._policy = SoftmaxPolicy(self, self.num_actions)
        self._baseline = StateValueFunction(self)
        self._bootstrap_values = Fetch("bootstrap_values", self, "bootstrap_values")
        self._engine = None
        self._engine_key = None
        self._pending_checkpoint = None   # checkpoint.Saver.restore before the first train step (a2c_acktr.py:100-102)
        checkpoint._register_model(self)
        self._layer_tokens = {n: (Fetch("inputs", self, n + "/inputs"), Fetch("outputs", self, n + "/outputs"))
                              for n in eng.LAYERS}
        # the two heads are registered with the SAME inputs tensor (envs/atari/model.py:243,246) -> one shared factor
        self._layer_tokens["fc_baseline"] = (self._layer_tokens["fc_policy"][0], self._layer_tokens["fc_baseline"][1])
        self.engine_options = {}     # extra EngineConfig fields (precision, num_locations_mode, world_size, seed ...)

    # ------------------------------------------------------------------ K-FAC registration
    def register_layers(self, layer_collection):
        """envs/atari/model.py:219-246: three conv2d blocks (strides 4/2/1, VALID) and three fully connected blocks."""
        for name, stride in (("conv1", 4), ("conv2", 2), ("conv3", 1)):
            inputs, outputs = self._layer_tokens[name]
            layer_collection.register_conv2d(params=(name + "/weights", name + "/bias"), strides=[1, stride, stride, 1],
                                             padding="VALID", inputs=inputs, outputs=outputs)
        for name in ("fc4", "fc_policy", "fc_baseline"):
            inputs, outputs = self._layer_tokens[name]
            layer_collection.register_fully_connected(params=(name + "/weights", name + "/bias"), inputs=inputs,
                                                      outputs=outputs)

    # ------------------------------------------------------------------ variables
    def get_variables(self):
        """dict name -> numpy array in the reference's shapes (HWIO kernels, [in, out] matrices)."""
        if self._engine is not None:
            self._params = self._engine.get_params()
        return {k: v.copy() for k, v in self._params.items()}

    def set_variables(self, params):
        shapes = eng.param_shapes(self.num_actions, self.conv3_num_filters)
        for k, shape in shapes.items():
            if tuple(np.shape(params[k])) != shape:
                raise ValueError("variable %s: expected shape %s, got %s" % (k, shape, np.shape(params[k])))
        self._params = {k: np.asarray(params[k], np.float32).copy() for k in shapes}
        if self._engine is not None:
            self._engine.set_params(self._params)

    @property
    def engine(self):
        return self._engine

    # ------------------------------------------------------------------ used by Session
    def _build_engine(self, session, num_envs, num_steps, objective):
        kw = dict(num_envs=num_envs, num_steps=num_steps, num_actions=self.num_actions,
                  conv3_filters=self.conv3_num_filters)
        if objective is not None:
            kw.update(gamma=objective.discount_factor, entropy_beta=objective.entropy_regularization_strength,
                      value_loss_weight=objective._baseline_loss_weight)
            if objective._optimizer is not None:
                kw.update(objective._optimizer.engine_overrides())
            elif getattr(objective, "_separate", None) is not None:
                kw.update(acktr=False)      # optimize_separate: first-order engine, the optimizers run on its gradients
        else:
            kw.update(acktr=False)
        if session.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            kw.setdefault("world_size", torch.distributed.get_world_size(session.group))
        if "seed" not in self.engine_options:
            # every rank of a data-parallel job (and every model without a fixed seed) gets its own Philox stream for action
            # sampling and Fisher sampling
            base = self.random_seed if self.random_seed is not None else int.from_bytes(__import__("os").urandom(4), "little")
            rank = (torch.distributed.get_rank(session.group)
                    if torch.distributed.is_available() and torch.distributed.is_initialized() else 0)
            kw["seed"] = (int(base) * 1000003 + rank) & 0x7FFFFFFFFFFFFFFF
        kw.update(self.engine_options)
        cfg = eng.EngineConfig(**kw)
        old = self._engine
        if old is not None:
            self._params = old.get_params()
        e = eng.Engine(cfg, session.device)
        e.set_params(self._params)
        if old is not None and old.config.acktr == cfg.acktr and old.num_params == e.num_params:
            sd = old.state_dict()
            e.load_state_dict(sd)        # the learner state does not depend on the batch shape
        if self._pending_checkpoint is not None and (objective is None or objective._optimizer is not None):
            pending = self._pending_checkpoint
            if objective is not None:
                self._pending_checkpoint = None      # an acting-only engine keeps it for the learner built later
            checkpoint.arrays_to_state(e, pending)
        e.is_learner = objective is not None
        self._engine = e
        self._engine_key = (num_envs, num_steps, id(objective) if objective is not None else None)
        if objective is not None and objective._global_step is not None:
            objective._global_step.bind(e)
        return e

    def _engine_for_feed(self, session, feed, objective):
        obs = feed.get(self._observations_placeholder)
        if obs is None:
            raise ValueError("observations_placeholder must be fed")
        shape = tuple(obs.shape) if hasattr(obs, "shape") else np.shape(obs)
        if len(shape) != 5 or tuple(shape[2:]) != eng.OBS_SHAPE:
            raise ValueError("observations must have shape [environments, steps, 84, 84, 4], got %s" % (shape,))
        key = (shape[0], shape[1], id(objective))
        if self._engine is None or self._engine_key != key:
            self._build_engine(session, shape[0], shape[1], objective)
        return self._engine

    def _train_feed(self, feed):
        names = ("observations", "bootstrap_observations", "actions", "rewards", "terminals")
        phs = (self._observations_placeholder, self._bootstrap_observations_placeholder, self._actions_placeholder,
               self._rewards_placeholder, self._terminals_placeholder)
        out = []
        for name, ph in zip(names, phs):
            if ph not in feed:
                raise ValueError("placeholder '%s' must be fed for a train step (a2c_acktr.py:117-126)" % name)
            v = feed[ph]
            if not isinstance(v, torch.Tensor):
                v = np.asarray(v, dtype=ph.dtype if ph.dtype != np.bool_ else np.bool_)
            out.append(v)
        return out

    def _act(self, session, feed, greedy):
        obs = feed.get(self._observations_placeholder)
        if obs is None:
            raise ValueError("observations_placeholder must be fed")
        if not isinstance(obs, torch.Tensor):
            obs = torch.from_numpy(np.ascontiguousarray(np.asarray(obs, np.uint8)))
        lead = tuple(obs.shape[:-3])
        flat = obs.reshape((-1,) + eng.OBS_SHAPE)
        e = self._engine
        if e is None or flat.shape[0] > e.rows + e.num_envs:
            e = self._build_engine(session, flat.shape[0], 1, None) if e is None else e
            if flat.shape[0] > e.rows + e.num_envs:
                raise ValueError("too many observations (%d) for the engine's batch (%d)" % (flat.shape[0], e.rows + e.num_envs))
        actions, logits, values = e.act(flat, greedy=greedy, want_logits=True)
        torch.cuda.current_stream(session.device).synchronize()
        a = actions.cpu().numpy()
        # DistributionPolicy.sample / mode squeeze the last (step) axis: valid for a step dimension of 1 (policies.py:86-87)
        shaped = a.reshape(lead)
        if len(lead) == 2 and lead[1] == 1:
            shaped = shaped.reshape(lead[0])
        z = logits.cpu().numpy().reshape(lead + (self.num_actions,))
        v = values.cpu().numpy().reshape(lead)
        return {"sample": shaped, "mode": shaped, "logits": z, "value": v}

    def _bootstrap_only(self, session, feed):
        obs = feed.get(self._bootstrap_observations_placeholder)
        if obs is None:
            raise ValueError("bootstrap_observations_placeholder must be fed")
        res = self._act(session, {self._observations_placeholder: obs}, greedy=True)
        return res["value"]
