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

    policy_loss = property(lambda self: self._policy_loss)
    baseline_loss = property(lambda self: self._baseline_loss)
    mean_entropy = property(lambda self: self._mean_entropy)

    def optimize_separate(self, policy_optimizer, baseline_optimizer, policy_kwargs=None, baseline_kwargs=None):
        """objectives.py:31-54.  Like the reference, the default None kwargs raise TypeError (`**None`);
        separate optimizers for policy and baseline are outside the accelerated path."""
        dict(**policy_kwargs)
        dict(**baseline_kwargs)
        raise NotImplementedError("optimize_separate is not on the ACKTR hot path; use optimize_shared")

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
