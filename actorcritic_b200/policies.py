"""Policies (actorcritic/policies.py): the categorical softmax policy of the hot path.  The attributes are fetch
tokens for `Session.run`; the arithmetic (log-softmax, entropy, inverse-CDF sampling, argmax) is in libacx
(layers.cu: loss_grad_kernel, sample_actions_kernel)."""
from abc import ABCMeta, abstractmethod

from .session import Fetch


class Policy(object, metaclass=ABCMeta):
    """policies.py:9-64."""

    @property
    @abstractmethod
    def sample(self):
        pass

    @property
    @abstractmethod
    def mode(self):
        pass

    @property
    @abstractmethod
    def entropy(self):
        pass

    @property
    @abstractmethod
    def log_prob(self):
        pass

    @abstractmethod
    def register_predictive_distribution(self, layer_collection, random_seed=None):
        pass


class DistributionPolicy(Policy):
    """policies.py:67-121: sample / mode squeeze the step axis (only valid for a step dimension of 1)."""

    def __init__(self, model, name=None):
        self.model = model
        self._sample = Fetch("sample", self, "policy/sample")
        self._mode = Fetch("mode", self, "policy/mode")
        self._entropy = Fetch("entropy", self, "policy/entropy")
        self._log_prob = Fetch("log_prob", self, "policy/log_prob")

    sample = property(lambda self: self._sample)
    mode = property(lambda self: self._mode)
    entropy = property(lambda self: self._entropy)
    log_prob = property(lambda self: self._log_prob)

    def register_predictive_distribution(self, layer_collection, random_seed=None):
        raise NotImplementedError()


class SoftmaxPolicy(DistributionPolicy):
    """policies.py:124-158: Categorical(logits)."""

    def __init__(self, model, num_actions, name=None):
        super().__init__(model, name)
        self.num_actions = num_actions
        self.logits = Fetch("logits", self, "policy/logits")

    def register_predictive_distribution(self, layer_collection, random_seed=None):
        """policies.py:146-158."""
        return layer_collection.register_categorical_predictive_distribution(logits=self.logits, seed=random_seed)
