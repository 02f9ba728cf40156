"""Base classes shared by the plan implementations (planner.py:359-420 of the reference) and
the initialisers a .cfg file can name (planner.py:140-155 resolves them with getattr on
jax.nn.initializers; jax does not exist here, so the names map to local objects)."""
from __future__ import annotations

from abc import ABCMeta, abstractmethod

import torch


class _Initializer:
    def __init__(self, name: str, **kwargs):
        self.name, self.kwargs = name, kwargs

    def __call__(self, shape, generator: torch.Generator) -> torch.Tensor:
        if self.name == "normal":
            return torch.randn(shape, generator=generator) * self.kwargs.get("stddev", 1e-2)
        if self.name == "uniform":
            return torch.rand(shape, generator=generator) * self.kwargs.get("scale", 1e-2)
        if self.name == "zeros":
            return torch.zeros(shape)
        if self.name == "ones":
            return torch.ones(shape)
        if self.name == "constant":
            return torch.full(shape, float(self.kwargs.get("value", 0.0)))
        raise NotImplementedError(f"initializer {self.name}")

    def __repr__(self):
        return f"{self.name}({self.kwargs})"


class initializers:     # noqa: N801  (mirrors jax.nn.initializers)
    @staticmethod
    def normal(stddev: float = 1e-2):
        return _Initializer("normal", stddev=stddev)

    @staticmethod
    def uniform(scale: float = 1e-2):
        return _Initializer("uniform", scale=scale)

    @staticmethod
    def constant(value: float = 0.0):
        return _Initializer("constant", value=value)

    zeros = _Initializer("zeros")
    ones = _Initializer("ones")


class JaxPlan(metaclass=ABCMeta):
    """Base class of policy representations (planner.py:359-420)."""
    history_dependent = False

    def __init__(self) -> None:
        self.bounds = None

    @abstractmethod
    def compile(self, compiled, test_compiled, _bounds, horizon, preprocessor=None) -> None:
        pass

    @abstractmethod
    def guess_next_epoch(self, params):
        pass
