"""Base class of actor-critic models - the train-step contract of the reference (actorcritic/model.py:11-186)
kept as the drop-in boundary: five placeholders, `policy`, `baseline`, `bootstrap_values`, K-FAC registration,
`sample_actions` / `select_max_actions`."""
from abc import ABCMeta

import numpy as np

from . import spaces
from .session import Fetch, Placeholder


class ActorCriticModel(object, metaclass=ABCMeta):
    def __init__(self, observation_space, action_space):
        self._observations_placeholder = None
        self._bootstrap_observations_placeholder = None
        self._actions_placeholder = None
        self._rewards_placeholder = None
        self._terminals_placeholder = None
        self._setup_placeholders(observation_space, action_space)
        self._policy = None
        self._baseline = None
        self._bootstrap_values = None

    observations_placeholder = property(lambda self: self._observations_placeholder)
    bootstrap_observations_placeholder = property(lambda self: self._bootstrap_observations_placeholder)
    actions_placeholder = property(lambda self: self._actions_placeholder)
    rewards_placeholder = property(lambda self: self._rewards_placeholder)
    terminals_placeholder = property(lambda self: self._terminals_placeholder)
    policy = property(lambda self: self._policy)
    baseline = property(lambda self: self._baseline)
    bootstrap_values = property(lambda self: self._bootstrap_values)

    def _setup_placeholders(self, observation_space, action_space):
        """model.py:97-105."""
        self._observations_placeholder = _space_placeholder(observation_space, [None, None], "observations")
        self._bootstrap_observations_placeholder = _space_placeholder(observation_space, [None], "bootstrap_observations")
        self._actions_placeholder = _space_placeholder(action_space, [None, None], "actions")
        self._rewards_placeholder = Placeholder("rewards", np.float32, [None, None])
        self._terminals_placeholder = Placeholder("terminals", np.bool_, [None, None])

    def register_layers(self, layer_collection):
        """model.py:107-119: models without K-FAC support raise NotImplementedError."""
        raise NotImplementedError()

    def register_predictive_distributions(self, layer_collection, random_seed=None):
        """model.py:121-133."""
        self._policy.register_predictive_distribution(layer_collection, random_seed)
        self._baseline.register_predictive_distribution(layer_collection, random_seed)

    # ------------------------------------------------------------------ generic acting path (models built from nn.*)
    def _forward_device(self, observations):
        """Hook for models assembled from `actorcritic_b200.nn` layers the way the reference assembles its graph
        (envs/atari/model.py:173-217): observations as a CUDA tensor [rows, *observation_shape] in the placeholder's dtype
        -> (logits fp32 [rows, num_actions], values fp32 [rows]).  AtariModel does not use it (its forward is fused inside
        the learner engine)."""
        raise NotImplementedError("%s defines no device forward" % type(self).__name__)

    def _act(self, session, feed, greedy):
        """sample / mode / logits / value of the fed observations through `_forward_device` and `acx_sample_actions`."""
        import torch
        from . import ops
        obs = feed.get(self._observations_placeholder)
        if obs is None:
            raise ValueError("observations_placeholder must be fed")
        if not isinstance(obs, torch.Tensor):
            obs = torch.from_numpy(np.ascontiguousarray(np.asarray(obs, self._observations_placeholder.dtype)))
        obs = obs.to(session.device, non_blocking=True)
        trailing = len(self._observations_placeholder.shape) - 2          # observation dimensions after [env, step]
        lead = tuple(obs.shape[: obs.dim() - trailing])
        flat = obs.reshape((-1,) + tuple(obs.shape[obs.dim() - trailing:]))
        logits, values = self._forward_device(flat)
        self._act_calls = getattr(self, "_act_calls", 0) + 1
        actions = ops.sample_actions(logits, seed=int(getattr(self, "random_seed", 0) or 0), step=self._act_calls, greedy=greedy)
        torch.cuda.current_stream(session.device).synchronize()
        shaped = actions.cpu().numpy().reshape(lead)
        if len(lead) == 2 and lead[1] == 1:      # DistributionPolicy.sample / mode squeeze the step axis (policies.py:86-87)
            shaped = shaped.reshape(lead[0])
        return {"sample": shaped, "mode": shaped, "logits": logits.cpu().numpy().reshape(lead + (logits.shape[-1],)),
                "value": values.cpu().numpy().reshape(lead)}

    def _bootstrap_only(self, session, feed):
        obs = feed.get(self._bootstrap_observations_placeholder)
        if obs is None:
            raise ValueError("bootstrap_observations_placeholder must be fed")
        return self._act(session, {self._observations_placeholder: obs}, greedy=True)["value"]

    def _engine_for_feed(self, session, feed, objective):
        raise NotImplementedError("%s has no fused learner engine: the train step of this framework is the ACKTR / A2C hot "
                                  "path of AtariModel (envs/atari/model.py)" % type(self).__name__)

    def sample_actions(self, observations, session):
        """model.py:135-151: returns a nested list shaped like the [environment, step] batch of `observations`."""
        return session.run(self.policy.sample, feed_dict={self.observations_placeholder: observations}).tolist()

    def select_max_actions(self, observations, session):
        """model.py:153-169."""
        return session.run(self.policy.mode, feed_dict={self.observations_placeholder: observations}).tolist()


def _space_placeholder(space, batch_shape=None, name=None):
    """model.py:172-186: Discrete(n) -> np.min_scalar_type(n) (uint8 for n <= 255); Box -> its dtype and shape."""
    if batch_shape is None:
        batch_shape = [None]
    if spaces.is_discrete(space):
        return Placeholder(name, np.min_scalar_type(space.n), batch_shape)
    if spaces.is_box(space):
        low, high = np.asarray(space.low), np.asarray(space.high)
        if low.dtype != high.dtype or low.shape != high.shape:
            raise TypeError()
        return Placeholder(name, low.dtype, list(batch_shape) + list(low.shape))
    raise TypeError("Unsupported space")
