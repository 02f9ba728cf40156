"""``JaxRDDLCompiler``: the exact-model compiler object of the reference API.

Mirror of pyRDDLGym_jax/core/compiler.py:78-167 for this build: it holds the lifted model,
the initial values, CPF levels and action bounds, and hands out rollout programs
(core/program.py) instead of JAX closures.  ``compile_rollouts`` (compiler.py:621-666) is
served by ``rollout_program``; the relaxed subclasses live in core/logic.py.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Sequence, Tuple

import numpy as np

from .lowering import ERR_BITS
from .program import PlanSpec, Program, ProgramBuilder


class JaxRDDLCompiler:
    """Compiles a lifted RDDL model into rollout programs with exact semantics."""

    relaxed = False

    ERROR_CODES = {"NORMAL": 0, **{
        ("INVALID_CAST" if k == "CAST" else
         "INVALID_PARAM_KRON_DELTA" if k == "KRON" else f"INVALID_PARAM_{k}"): 1 << v
        for k, v in ERR_BITS.items()}}

    INVERSE_ERROR_CODES = {     # compiler.py:890-916
        0: "Casting occurred that could result in loss of precision.",
        1: "Found Uniform(a, b) distribution where a > b.",
        2: "Found Normal(m, v^2) distribution where v < 0.",
        3: "Found Exponential(s) distribution where s <= 0.",
        4: "Found Weibull(k, l) distribution where either k <= 0 or l <= 0.",
        5: "Found Bernoulli(p) distribution where either p < 0 or p > 1.",
        6: "Found Poisson(l) distribution where l < 0.",
        7: "Found Gamma(k, l) distribution where either k <= 0 or l <= 0.",
        8: "Found Beta(a, b) distribution where either a <= 0 or b <= 0.",
        9: "Found Geometric(p) distribution where either p < 0 or p > 1.",
        10: "Found Pareto(k, l) distribution where either k <= 0 or l <= 0.",
        11: "Found Student(df) distribution where df <= 0.",
        12: "Found Gumbel(m, s) distribution where s <= 0.",
        13: "Found Laplace(m, s) distribution where s <= 0.",
        14: "Found Cauchy(m, s) distribution where s <= 0.",
        15: "Found Gompertz(k, l) distribution where either k <= 0 or l <= 0.",
        16: "Found ChiSquare(df) distribution where df <= 0.",
        17: "Found Kumaraswamy(a, b) distribution where either a <= 0 or b <= 0.",
        18: "Found Discrete(p) distribution where either p < 0 or p does not sum to 1.",
        19: "Found KronDelta(x) distribution where x is not int nor bool.",
    }

    def __init__(self, rddl, *args, allow_synchronous_state: bool = True,
                 logger: Optional[Any] = None, use64bit: bool = False,
                 stochastic_is_fluent: bool = False,
                 python_functions: Optional[Dict[str, Callable]] = None, **kwargs) -> None:
        if args:
            print(f"[WARN] JaxRDDLCompiler received invalid args {args}.")
        if kwargs:
            print(f"[WARN] JaxRDDLCompiler received invalid kwargs {kwargs}.")
        if use64bit:
            raise NotImplementedError("the B200 rollout kernels compute in fp32/int32 only")
        if python_functions:
            raise NotImplementedError("external Python functions cannot be lowered to the GPU")
        self.rddl = rddl
        self.logger = logger
        self.use64bit = False
        self.allow_synchronous_state = allow_synchronous_state
        self.python_functions = python_functions or {}
        self.REAL, self.INT = np.float32, np.int32
        self.JAX_TYPES = {"int": np.int32, "real": np.float32, "bool": np.bool_}
        self.init_values = {k: np.asarray(v) for k, v in rddl.init_values.items()}
        self.levels = rddl.levels
        self.print_warnings = True
        self._builders: Dict[Any, ProgramBuilder] = {}

    def get_kwargs(self) -> Dict[str, Any]:
        return {"allow_synchronous_state": self.allow_synchronous_state,
                "use64bit": self.use64bit, "python_functions": self.python_functions}

    def relax_config(self) -> Optional[Dict[str, Any]]:
        return None

    def compile(self, *args, **kwargs) -> None:
        """Kept for API compatibility (compiler.py:188): programs are built on demand."""
        self.model_aux = {"params": {}, "overriden": {}, "exact": set()}

    # ---------------------------------------------------------------------------------
    @property
    def bounds(self) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
        return self.rddl.bounds

    def builder(self, spec: PlanSpec, train_policy: bool, log_fluents: Sequence[str] = (),
                batch_hint: int = 0) -> ProgramBuilder:
        return ProgramBuilder(self.rddl, self.relax_config(), spec, train_policy, log_fluents,
                              batch_hint)

    @staticmethod
    def get_error_codes(error: int):
        return [i for i, c in enumerate(reversed(bin(int(error))[2:])) if c == "1"]

    @staticmethod
    def get_error_messages(error: int):
        return [JaxRDDLCompiler.INVERSE_ERROR_CODES.get(i, f"error bit {i}")
                for i in JaxRDDLCompiler.get_error_codes(error)]
