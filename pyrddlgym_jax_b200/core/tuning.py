"""Bayesian hyper-parameter search for the planner (the reference's ``JaxParameterTuning``,
core/tuning.py:54-583).

The reference drives the third-party ``bayes_opt`` package over a pool of worker processes (one planner per
worker, tuning.py:502-510).  Here the search is a self-contained Gaussian-process upper-confidence-bound loop
(scikit-learn's ``GaussianProcessRegressor`` on the unit cube, annealed exploration weight), and the trials of
an iteration run one after another on this process's GPU -- a planner run already fills the device; with
several GPUs, launch one tuner per GPU on a slice of the iterations.  The objective is the reference's: plug the
suggested values into a ``.cfg`` TEMPLATE (tags such as ``LEARNING_RATE_TUNE`` replaced textually,
tuning.py:207-213), train, act in the simulator and average the returns (tuning.py:241-336).
"""
from __future__ import annotations

import csv
import math
import time
from typing import Any, Callable, Dict, Iterable, List, Optional

import numpy as np

Kwargs = Dict[str, Any]
ParameterValues = Dict[str, Any]
COLUMNS = ["pid", "worker", "iteration", "target", "best_target", "acq_params"]


class Hyperparameter:
    """A tunable quantity: the optimiser searches ``[lower_bound, upper_bound]`` and ``search_to_config_map``
    turns the searched value into what is written in place of ``tag`` (tuning.py:54-67)."""

    def __init__(self, tag: str, lower_bound: float, upper_bound: float,
                 search_to_config_map: Callable) -> None:
        self.tag = tag
        self.lower_bound = float(lower_bound)
        self.upper_bound = float(upper_bound)
        self.search_to_config_map = search_to_config_map

    def __str__(self) -> str:
        name = getattr(self.search_to_config_map, "__name__", "map")
        return f"{name} : [{self.lower_bound}, {self.upper_bound}] -> {self.tag}"


def power_10(x):
    return 10.0 ** x


def power_2_int(x):
    return int(2 ** int(round(x)))


class UpperConfidenceBound:
    """mean + kappa * deviation with a geometric decay of kappa after ``exploration_decay_delay`` suggestions."""

    def __init__(self, kappa: float = 2.576, exploration_decay: Optional[float] = None,
                 exploration_decay_delay: int = 0) -> None:
        self.kappa = float(kappa)
        self.exploration_decay = exploration_decay
        self.exploration_decay_delay = int(exploration_decay_delay)
        self.i = 0

    def __call__(self, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
        return mean + self.kappa * std

    def step(self) -> None:
        self.i += 1
        if self.exploration_decay is not None and self.i > self.exploration_decay_delay:
            self.kappa *= self.exploration_decay

    def __str__(self) -> str:
        return f"UpperConfidenceBound(kappa={self.kappa:.4g}, decay={self.exploration_decay})"


class JaxParameterTuning:
    """Tunes the hyper-parameters of a planner configuration on one model (tuning.py:76)."""

    def __init__(self, env, config_template: str, hyperparams: Iterable[Hyperparameter], online: bool,
                 eval_trials: int = 5, rollouts_per_trial: int = 1, verbose: bool = True,
                 timeout_tuning: float = np.inf, pool_context: str = "spawn", num_workers: int = 1,
                 poll_frequency: float = 0.2, gp_iters: int = 25,
                 acquisition: Optional[UpperConfidenceBound] = None,
                 gp_init_kwargs: Optional[Kwargs] = None, gp_params: Optional[Kwargs] = None,
                 objective: Optional[Callable[[str, int], float]] = None) -> None:
        """``env``: the lifted RDDL model (or an object with a ``.model``).  ``objective``: replaces the
        train-and-act objective by ``objective(config_string, seed) -> return`` (tests, custom metrics)."""
        self.model = getattr(env, "model", env)
        self.env = env
        self.config_template = config_template
        self.hyperparams_dict = {h.tag: h for h in hyperparams if h.tag in config_template}
        self.online = bool(online)
        self.eval_trials = int(eval_trials)
        self.rollouts_per_trial = int(rollouts_per_trial)
        self.verbose = verbose
        self.timeout_tuning = timeout_tuning
        self.pool_context = pool_context
        self.num_workers = int(num_workers)
        self.poll_frequency = poll_frequency
        self.gp_iters = int(gp_iters)
        self.gp_init_kwargs = dict(gp_init_kwargs or {})
        self.gp_params = {"n_restarts_optimizer": 10, **(gp_params or {})}
        total = self.gp_iters * self.num_workers
        self.acquisition = acquisition if acquisition is not None else \
            self.annealing_acquisition(total, total // 4)
        self._objective = objective
        self.best_params: ParameterValues = {}
        self.best_target = -math.inf
        self.log: List[Dict[str, Any]] = []

    # ---- the reference's helpers ---------------------------------------------------------------------
    @staticmethod
    def make_default_kernel():
        """Three Matern-5/2 components of short, medium and long length scale with free weights (the surrogate
        the reference configures, tuning.py:152-161)."""
        from sklearn.gaussian_process.kernels import ConstantKernel, Matern
        parts = [(0.5, (0.1, 0.5)), (1.0, (0.5, 1.0)), (5.0, (1.0, 5.0))]
        kernel = None
        for scale, bounds in parts:
            term = ConstantKernel(1.0, (0.01, 100.0)) * Matern(length_scale=scale, length_scale_bounds=bounds,
                                                               nu=2.5)
            kernel = term if kernel is None else kernel + term
        return kernel

    @staticmethod
    def annealing_acquisition(n_samples: int, n_delay_samples: int = 0, kappa1: float = 10.0,
                              kappa2: float = 1.0) -> UpperConfidenceBound:
        """Exploration weight decaying geometrically from kappa1 to kappa2 over the run (tuning.py:184-193)."""
        span = max(1, n_samples - n_delay_samples)
        return UpperConfidenceBound(kappa=kappa1, exploration_decay=(kappa2 / kappa1) ** (1.0 / span),
                                    exploration_decay_delay=n_delay_samples)

    @staticmethod
    def search_to_config_params(hyperparams: Dict[str, Hyperparameter],
                                params: ParameterValues) -> ParameterValues:
        return {tag: h.search_to_config_map(params[tag]) for tag, h in hyperparams.items()}

    @staticmethod
    def config_from_template(config_template: str, config_params: ParameterValues) -> str:
        out = config_template
        for tag, value in config_params.items():
            out = out.replace(tag, str(value))
        return out

    def summarize_hyperparameters(self) -> str:
        table = "\n".join(f"        {h}" for h in self.hyperparams_dict.values())
        return (f"hyperparameter optimizer parameters:\n"
                f"    tuned_hyper_parameters    =\n{table}\n"
                f"    gp_params                 ={self.gp_params}\n"
                f"    tuning_iterations         ={self.gp_iters}\n"
                f"    tuning_timeout            ={self.timeout_tuning}\n"
                f"    tuning_batch_size         ={self.num_workers}\n"
                f"meta-objective parameters:\n"
                f"    planning_trials_per_iter  ={self.eval_trials}\n"
                f"    rollouts_per_trial        ={self.rollouts_per_trial}\n"
                f"    acquisition_fn            ={self.acquisition}")

    @property
    def best_config(self) -> str:
        return self.config_from_template(
            self.config_template, self.search_to_config_params(self.hyperparams_dict, self.best_params))

    def build_best_controller(self):
        from .planner import (JaxBackpropPlanner, JaxOfflineController, JaxOnlineController,
                              load_config_from_string)
        planner_args, _, train_args = load_config_from_string(self.best_config)
        planner_args.pop("dashboard", None)
        planner = JaxBackpropPlanner(rddl=self.model, **planner_args)
        return (JaxOnlineController if self.online else JaxOfflineController)(planner, **train_args)

    # ---- the meta-objective (tuning.py:241-391) ------------------------------------------------------
    def evaluate_config(self, config_string: str, seed: int) -> float:
        """Average return of ``eval_trials`` independent trainings of the configuration, each followed by
        ``rollouts_per_trial`` episodes in the simulator (offline), or of ``eval_trials`` episodes of online
        replanning."""
        if self._objective is not None:
            return float(self._objective(config_string, seed))
        from ..entry_point import evaluate
        from .planner import (JaxBackpropPlanner, JaxOfflineController, JaxOnlineController,
                              load_config_from_string)
        from .simulator import JaxRDDLSimulator
        planner_args, _, train_args = load_config_from_string(config_string)
        planner_args.pop("dashboard", None)
        planner_args.pop("parallel_updates", None)
        train_args = {**train_args, "print_summary": False, "print_progress": False}
        planner = JaxBackpropPlanner(rddl=self.model, **planner_args)
        sim = JaxRDDLSimulator(self.model, key=seed, keep_tensors=True)
        total = 0.0
        for trial in range(self.eval_trials):
            key = seed * 1000 + trial
            if self.online:
                agent = JaxOnlineController(planner, key=key, **{k: v for k, v in train_args.items()
                                                                 if k != "key"})
                returns = evaluate(agent, sim, episodes=1, verbose=False)
            else:
                agent = JaxOfflineController(planner, key=key, train_on_reset=False,
                                             **{k: v for k, v in train_args.items() if k != "key"})
                returns = evaluate(agent, sim, episodes=self.rollouts_per_trial, verbose=False)
            total += float(np.mean(returns))
        return total / self.eval_trials

    # ---- the search ------------------------------------------------------------------------------------
    def _suggest(self, X: np.ndarray, y: np.ndarray, rng: np.random.Generator, batch: int) -> np.ndarray:
        d = len(self.hyperparams_dict)
        if len(y) < max(2, d):                              # not enough points for a surrogate yet
            return rng.random((batch, d))
        from sklearn.gaussian_process import GaussianProcessRegressor
        gp = GaussianProcessRegressor(kernel=self.make_default_kernel(), normalize_y=True, alpha=1e-6,
                                      random_state=int(rng.integers(2 ** 31)), **self.gp_params)
        import warnings
        with warnings.catch_warnings():                     # (lbfgs convergence notes of the kernel fit)
            warnings.simplefilter("ignore")
            gp.fit(X, y)
        self._gp = gp
        picks = []
        for _ in range(batch):
            cand = rng.random((2048, d))
            if picks:                                       # spread a batch: penalise the points already taken
                dist = np.min(np.linalg.norm(cand[:, None, :] - np.asarray(picks)[None], axis=2), axis=1)
            else:
                dist = np.ones(len(cand))
            mean, std = gp.predict(cand, return_std=True)
            picks.append(cand[int(np.argmax(self.acquisition(mean, std * np.minimum(1.0, 4.0 * dist))))])
        return np.asarray(picks)

    def tune(self, key: int, log_file: Optional[str] = None, show_dashboard: bool = False,
             print_hyperparams: bool = False) -> ParameterValues:
        """Runs ``gp_iters`` iterations of ``num_workers`` suggestions each (or until ``timeout_tuning``);
        returns the best search-space values found (``best_config`` is the concrete configuration)."""
        if print_hyperparams or self.verbose:
            print(self.summarize_hyperparameters())
        tags = list(self.hyperparams_dict)
        lo = np.array([self.hyperparams_dict[t].lower_bound for t in tags])
        hi = np.array([self.hyperparams_dict[t].upper_bound for t in tags])
        rng = np.random.default_rng(int(key))
        X, y = np.zeros((0, len(tags))), np.zeros(0)
        writer = handle = None
        if log_file is not None:
            handle = open(log_file, "w", newline="")
            writer = csv.writer(handle)
            writer.writerow(COLUMNS + tags)
        start = time.time()
        try:
            for it in range(self.gp_iters):
                if time.time() - start > self.timeout_tuning:
                    break
                for w, u in enumerate(self._suggest(X, y, rng, self.num_workers)):
                    params = dict(zip(tags, (lo + u * (hi - lo)).tolist()))
                    config_params = self.search_to_config_params(self.hyperparams_dict, params)
                    if self.verbose:
                        print(f"[{it}:{w}] " + ", ".join(f"{k}={v}" for k, v in config_params.items()),
                              flush=True)
                    target = self.evaluate_config(
                        self.config_from_template(self.config_template, config_params),
                        int(rng.integers(2 ** 20)))
                    if not np.isfinite(target):
                        target = float(np.min(y)) if len(y) else -1e30
                    X, y = np.vstack([X, u[None]]), np.append(y, target)
                    if target > self.best_target:
                        self.best_target, self.best_params = target, params
                    self.acquisition.step()
                    row = {"pid": 0, "worker": w, "iteration": it, "target": target,
                           "best_target": self.best_target, "acq_params": str(self.acquisition), **params}
                    self.log.append(row)
                    if writer is not None:
                        writer.writerow([row[c] for c in COLUMNS] + [params[t] for t in tags])
                        handle.flush()
                    if self.verbose:
                        print(f"[{it}:{w}] target={target:.5f} best={self.best_target:.5f}", flush=True)
        finally:
            if handle is not None:
                handle.close()
        return self.best_params
