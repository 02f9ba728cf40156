"""Hyper-parameter search (core/tuning.py; reference core/tuning.py:54-583): template substitution, search
maps and the Gaussian-process UCB loop on a synthetic objective (the train-and-act objective needs a GPU:
tests/test_simulator_gpu.py::test_tuning_on_a_small_budget)."""
import csv
import math
import re

import numpy as np

from pyrddlgym_jax_b200.core.tuning import Hyperparameter, JaxParameterTuning, power_10, power_2_int

TEMPLATE = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
sigmoid_weight=MODEL_WEIGHT_TUNE

[Planner]
method='JaxStraightLinePlan'
method_kwargs={}
optimizer='rmsprop'
optimizer_kwargs={'learning_rate': LEARNING_RATE_TUNE}
batch_size_train=BATCH_TUNE
batch_size_test=BATCH_TUNE

[Optimize]
train_seconds=1
print_summary=False
print_progress=False
"""


def _params():
    return [Hyperparameter("MODEL_WEIGHT_TUNE", -1.0, 4.0, power_10),
            Hyperparameter("LEARNING_RATE_TUNE", -5.0, 1.0, power_10),
            Hyperparameter("BATCH_TUNE", 3.0, 8.0, power_2_int),
            Hyperparameter("NOT_IN_TEMPLATE", 0.0, 1.0, float)]


def test_template_substitution_and_maps():
    tuner = JaxParameterTuning(None, TEMPLATE, _params(), online=False, gp_iters=1, verbose=False,
                               objective=lambda cfg, seed: 0.0)
    assert set(tuner.hyperparams_dict) == {"MODEL_WEIGHT_TUNE", "LEARNING_RATE_TUNE", "BATCH_TUNE"}
    vals = JaxParameterTuning.search_to_config_params(
        tuner.hyperparams_dict, {"MODEL_WEIGHT_TUNE": 2.0, "LEARNING_RATE_TUNE": -3.0, "BATCH_TUNE": 5.4})
    assert vals == {"MODEL_WEIGHT_TUNE": 100.0, "LEARNING_RATE_TUNE": 0.001, "BATCH_TUNE": 32}
    cfg = JaxParameterTuning.config_from_template(TEMPLATE, vals)
    assert "sigmoid_weight=100.0" in cfg and "'learning_rate': 0.001" in cfg and "batch_size_train=32" in cfg
    from pyrddlgym_jax_b200.core.planner import load_config_from_string
    planner_args, _, train_args = load_config_from_string(cfg)
    assert planner_args["batch_size_train"] == 32 and planner_args["compiler_kwargs"]["sigmoid_weight"] == 100.0
    assert "MODEL_WEIGHT_TUNE" in tuner.summarize_hyperparameters()
    acq = JaxParameterTuning.annealing_acquisition(20, 5, kappa1=10.0, kappa2=1.0)
    for _ in range(20):
        acq.step()
    assert abs(acq.kappa - 1.0) < 1e-9


def test_gp_ucb_search_finds_the_optimum(tmp_path):
    def objective(cfg, seed):
        w = float(re.search(r"sigmoid_weight=([0-9.e+-]+)", cfg).group(1))
        lr = float(re.search(r"'learning_rate': ([0-9.e+-]+)", cfg).group(1))
        noise = 0.02 * math.sin(seed)
        return -(math.log10(lr) + 2.0) ** 2 - 0.5 * (math.log10(w) - 1.0) ** 2 + noise
    log = tmp_path / "tune.csv"
    tuner = JaxParameterTuning(None, TEMPLATE.replace("BATCH_TUNE", "32"), _params(), online=False,
                               gp_iters=14, num_workers=2, verbose=False, objective=objective)
    best = tuner.tune(key=3, log_file=str(log))
    assert abs(best["LEARNING_RATE_TUNE"] + 2.0) < 0.6 and abs(best["MODEL_WEIGHT_TUNE"] - 1.0) < 1.0
    assert tuner.best_target > -0.5
    rows = list(csv.reader(open(log)))
    assert rows[0][:6] == ["pid", "worker", "iteration", "target", "best_target", "acq_params"]
    assert len(rows) == 1 + 14 * 2
    bests = [float(r[4]) for r in rows[1:]]
    assert all(b >= a for a, b in zip(bests, bests[1:]))
    # the surrogate phase beats pure random search with the same budget on this objective
    rng = np.random.default_rng(3)
    rand_best = max(objective(TEMPLATE.replace("MODEL_WEIGHT_TUNE", str(10 ** (5 * u - 1)))
                              .replace("LEARNING_RATE_TUNE", str(10 ** (6 * v - 5))), 0)
                    for u, v in rng.random((28, 2)))
    assert tuner.best_target >= rand_best - 0.3
    assert "learning_rate" in tuner.best_config and "LEARNING_RATE_TUNE" not in tuner.best_config
