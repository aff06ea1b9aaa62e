"""Baselines (actorcritic/baselines.py): the state-value head."""
from abc import ABCMeta, abstractmethod

from .session import Fetch


class Baseline(object, metaclass=ABCMeta):
    """baselines.py:6-33."""

    @property
    @abstractmethod
    def value(self):
        pass

    @abstractmethod
    def register_predictive_distribution(self, layer_collection, random_seed=None):
        pass


class StateValueFunction(Baseline):
    """baselines.py:36-69: the Fisher of the value head is that of a unit-variance normal around the value."""

    def __init__(self, model, name=None):
        self.model = model
        self._value = Fetch("value", self, "baseline/value")

    value = property(lambda self: self._value)

    def register_predictive_distribution(self, layer_collection, random_seed=None):
        """baselines.py:55-66."""
        return layer_collection.register_normal_predictive_distribution(mean=self._value, var=1.0, seed=random_seed)
