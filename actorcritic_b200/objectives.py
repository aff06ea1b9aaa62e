"""actorcritic/objectives.py: the A2C objective and `optimize_shared`.  The loss values and gradients are computed by
libacx (returns.cu, layers.cu loss_grad_kernel); these classes carry the hyper-parameters and hand out fetch tokens."""
from abc import ABCMeta

from .session import Fetch


class ActorCriticObjective(object, metaclass=ABCMeta):
    """objectives.py:10-79."""

    def __init__(self, model):
        self.model = model
        self._policy_loss = Fetch("policy_loss", self)
        self._baseline_loss = Fetch("baseline_loss", self)
        self._mean_entropy = Fetch("mean_entropy", self)
        self._optimizer = None
        self._baseline_loss_weight = 0.5
        self._global_step = None
        self._separate = None

    policy_loss = property(lambda self: self._policy_loss)
    baseline_loss = property(lambda self: self._baseline_loss)
    mean_entropy = property(lambda self: self._mean_entropy)

    def optimize_separate(self, policy_optimizer, baseline_optimizer, policy_kwargs=None, baseline_kwargs=None):
        """objectives.py:31-54: `policy_optimizer.minimize(policy_loss, **policy_kwargs)` and
        `baseline_optimizer.minimize(baseline_loss, **baseline_kwargs)` grouped into one op - two backward passes from the
        same parameters (acx_learner_set_loss_weights (1, 0) and (0, 1)), each applied with its own optimizer slots
        (acx_clip_rmsprop_step / acx_clip_momentum_step).  Like the reference, the default None kwargs raise TypeError
        (`**None`).  Optimizers: [ClipGlobalNormOptimizer](RMSPropOptimizer | MomentumOptimizer)."""
        from . import nn
        policy_kwargs = dict(**policy_kwargs)
        baseline_kwargs = dict(**baseline_kwargs)
        self._separate = dict(policy=nn.standalone_spec(policy_optimizer), baseline=nn.standalone_spec(baseline_optimizer),
                              policy_global_step=policy_kwargs.get("global_step"),
                              baseline_global_step=baseline_kwargs.get("global_step"))
        self._optimizer = None
        self._global_step = policy_kwargs.get("global_step") or baseline_kwargs.get("global_step")
        return Fetch("optimize_separate", self, "optimize_op")

    def optimize_shared(self, optimizer, baseline_loss_weight=0.5, **kwargs):
        """objectives.py:56-79: minimise policy_loss + baseline_loss_weight * baseline_loss with one optimizer.
        kwargs: `global_step` (a2c_acktr.py:76).  Returns the optimize-op token for `Session.run`."""
        if not hasattr(optimizer, "engine_overrides"):
            raise TypeError("unsupported optimizer %r" % (optimizer,))
        self._optimizer = optimizer
        self._baseline_loss_weight = float(baseline_loss_weight)
        self._global_step = kwargs.get("global_step")
        return Fetch("optimize", self, "optimize_op")


class A2CObjective(ActorCriticObjective):
    """objectives.py:82-175."""

    def __init__(self, model, discount_factor=0.99, entropy_regularization_strength=0.01, name=None):
        super().__init__(model)
        self.discount_factor = float(discount_factor)
        self.entropy_regularization_strength = float(entropy_regularization_strength)
