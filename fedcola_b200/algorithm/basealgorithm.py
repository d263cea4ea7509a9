"""Mirror of /root/reference/src/algorithm/basealgorithm.py:5-15."""
from abc import ABCMeta, abstractmethod


class BaseOptimizer(metaclass=ABCMeta):
    """Federated optimization algorithm."""

    @abstractmethod
    def step(self, closure=None):
        raise NotImplementedError

    @abstractmethod
    def accumulate(self, **kwargs):
        raise NotImplementedError
