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


class Preprocessor(metaclass=ABCMeta):
    """Base class of state preprocessors (planner.py:255-281): initialize() -> stats,
    update(fls, stats) -> stats, transform(fls, stats) -> fls."""

    HYPERPARAMS_KEY = "__preprocessor__"

    def __init__(self) -> None:
        self._stats = None

    @abstractmethod
    def compile(self, compiled) -> None:
        pass

    def initialize(self):
        return self._stats

    def update(self, fls, stats):
        return stats


class StaticNormalizer(Preprocessor):
    """Min-max scaling of the non-boolean observed fluents by their box constraints
    (planner.py:284-338): the bounds come from the domain's state-invariants /
    action-preconditions, user bounds override them; fluents without a finite, non-degenerate box
    pass through."""

    def __init__(self, fluent_bounds=None) -> None:
        super().__init__()
        self.fluent_bounds = dict(fluent_bounds or {})

    def compile(self, compiled) -> None:
        import numpy as np
        rddl = compiled.rddl
        observed = rddl.observ_fluents if rddl.observ_fluents else rddl.state_fluents
        stats = {}
        for var in observed:
            if rddl.variable_ranges[var] == "bool":
                continue
            lower, upper = rddl.state_bounds.get(var, (-np.inf, np.inf))
            if np.all(np.isfinite(lower) & np.isfinite(upper) & np.less(lower, upper)):
                stats[var] = (np.asarray(lower, np.float32), np.asarray(upper, np.float32))
            user = self.fluent_bounds.get(var)
            if user is not None:
                stats[var] = tuple(np.asarray(b, np.float32) for b in user)
        self._stats = stats

    def transform(self, fls, stats):
        """(values - lower) / (upper - lower) on the fluents that have stats (torch or numpy)."""
        out = {}
        for var, values in fls.items():
            if var in stats:
                lower, upper = (torch.as_tensor(b, dtype=torch.float32) for b in stats[var])
                v = torch.as_tensor(values, dtype=torch.float32)
                out[var] = (v - lower) / (upper - lower)
            else:
                out[var] = values
        return out


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
