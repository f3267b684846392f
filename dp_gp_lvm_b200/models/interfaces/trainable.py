"""Abstract base class for a trainable model (reference src/models/interfaces/trainable.py:8-22)."""
from abc import ABC, abstractmethod


class Trainable(ABC):
    @property
    @abstractmethod
    def objective(self):
        """The cost to minimise."""
        pass
