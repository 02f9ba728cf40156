"""Command line of the planner (the reference's ``jaxplan`` entry point, entry_point.py:6-57, and
examples/run_plan.py:25-61):

    python -m pyrddlgym_jax_b200 plan <domain> <instance> <method> [-e EPISODES] [--train-seconds S]

<domain> / <instance>: RDDL file paths, or the name of a bundled fixture directory under
``pyrddlgym_jax_b200/domains`` and of an instance file in it (rddlrepository is not part of this build).
<method>: ``slp`` / ``drp`` (offline: optimise once, then act) or ``replan`` (online: re-optimise at every
decision epoch), or the path of a ``.cfg`` file in the reference's format.  The environment is the
``JaxRDDLSimulator`` of this package (one kernel launch per step).

    python -m pyrddlgym_jax_b200 tune <domain> <instance> <method> [-t TRIALS] [-i ITERS] [-w WORKERS] [-f FILE]

runs the Bayesian hyper-parameter search of ``core/tuning.py`` on the reference's tuning templates.
"""
from __future__ import annotations

import argparse
import os
from typing import Optional

# the settings of the reference's default_slp / default_drp / default_replan configuration files
_DEFAULT_CFG = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'

[Planner]
method='{plan}'
method_kwargs={{}}
optimizer='rmsprop'
optimizer_kwargs={{'learning_rate': {lr}}}
batch_size_train=32
batch_size_test=32
{extra}
[Optimize]
key=42
epochs={epochs}
train_seconds={seconds}
{optimize_extra}
"""
DEFAULTS = {
    "slp": dict(plan="JaxStraightLinePlan", lr=0.01, extra="", epochs=30000, seconds=60, optimize_extra=""),
    "drp": dict(plan="JaxDeepReactivePolicy", lr=0.0001, extra="", epochs=30000, seconds=60,
                optimize_extra=""),
    "replan": dict(plan="JaxStraightLinePlan", lr=0.01, extra="rollout_horizon=5\n", epochs=2000, seconds=1,
                   optimize_extra="print_summary=False"),
}
DOMAINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "domains")


def resolve_model_files(domain: str, instance: str):
    """(domain file, instance file) from paths or from the names of a bundled fixture."""
    if os.path.isfile(domain):
        return domain, (instance if os.path.isfile(instance) else None)
    d = os.path.join(DOMAINS, domain)
    if not os.path.isdir(d):
        raise FileNotFoundError(f"<{domain}> is neither an RDDL file nor one of {sorted(os.listdir(DOMAINS))}")
    for cand in (instance, f"instance{instance}", f"{instance}.rddl", f"instance{instance}.rddl"):
        path = cand if os.path.isfile(cand) else os.path.join(d, cand if cand.endswith(".rddl")
                                                              else cand + ".rddl")
        if os.path.isfile(path):
            return os.path.join(d, "domain.rddl"), path
    raise FileNotFoundError(f"no instance <{instance}> in {d}")


def load_method(method: str, train_seconds: Optional[float] = None):
    """(planner_args, plan_kwargs, train_args, online) of a method name or a .cfg path."""
    from .core.planner import load_config, load_config_from_string
    if method in DEFAULTS:
        planner_args, plan_kwargs, train_args = load_config_from_string(_DEFAULT_CFG.format(**DEFAULTS[method]))
    elif os.path.isfile(method):
        planner_args, plan_kwargs, train_args = load_config(method)
    else:
        raise ValueError("method must be slp, drp, replan, or a path to a valid .cfg file.")
    if train_seconds is not None:
        train_args["train_seconds"] = float(train_seconds)
    return planner_args, plan_kwargs, train_args, method == "replan"


def evaluate(controller, sim, episodes: int = 1, horizon: Optional[int] = None, verbose: bool = True):
    """The loop of pyRDDLGym's ``BaseAgent.evaluate`` (examples/run_plan.py:61): act in the simulator for
    ``episodes`` episodes; returns the list of total rewards."""
    horizon = sim.rddl.horizon if horizon is None else horizon
    gamma = float(sim.rddl.discount)
    totals = []
    for ep in range(episodes):
        state, done = sim.reset()
        controller.reset()
        total, disc = 0.0, 1.0
        for t in range(horizon):
            action = controller.sample_action(state)
            state, reward, done = sim.step(action)
            total += disc * float(reward)
            disc *= gamma
            if verbose:
                print(f"episode {ep} step {t}: reward {float(reward):.5f}")
            if done:
                break
        totals.append(total)
        if verbose:
            print(f"episode {ep} ended with return {total:.5f}")
    return totals


def run_plan(domain: str, instance: str, method: str, episodes: int = 1,
             train_seconds: Optional[float] = None, verbose: bool = True):
    from .core.planner import JaxBackpropPlanner, JaxOfflineController, JaxOnlineController
    from .core.simulator import JaxRDDLSimulator
    from .rddl.model import load_model
    dfile, ifile = resolve_model_files(domain, instance)
    model = load_model(dfile, ifile)
    planner_args, _, train_args, online = load_method(method, train_seconds)
    planner_args.pop("dashboard", None)
    planner = JaxBackpropPlanner(rddl=model, **planner_args)
    controller = (JaxOnlineController if online else JaxOfflineController)(planner, **train_args)
    sim = JaxRDDLSimulator(model, key=train_args.get("key", 42), keep_tensors=True)
    totals = evaluate(controller, sim, episodes=episodes, verbose=verbose)
    if verbose:
        import numpy as np
        print(f"mean return over {episodes} episode(s): {np.mean(totals):.5f}")
    return totals


# the tuning templates of the reference (examples/configs/tuning_{slp,drp,replan}.cfg): relaxation weight and
# learning rate (and, for drp, the layer width; for replan, the lookahead) are the tuned tags
_TUNE_CFG = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
sigmoid_weight=MODEL_WEIGHT_TUNE
print_warnings=False

[Planner]
method='{plan}'
method_kwargs={method_kwargs}
optimizer='rmsprop'
optimizer_kwargs={{'learning_rate': LEARNING_RATE_TUNE}}
batch_size_train=32
batch_size_test=32
pgpe=None
{extra}
[Optimize]
train_seconds={seconds}
print_summary=False
print_progress=False
"""


def run_tune(domain: str, instance: str, method: str, trials: int = 5, iters: int = 20, workers: int = 4,
             filepath: str = "", train_seconds: float = 30.0):
    """examples/run_tune.py of the reference: tune the relaxation weight and the learning rate (and the layer
    width of a deep policy / the lookahead of online replanning)."""
    from .core.tuning import Hyperparameter, JaxParameterTuning, power_10, power_2_int
    from .rddl.model import load_model
    if method not in ("slp", "drp", "replan"):
        raise ValueError("method must be slp, drp or replan")
    model = load_model(*resolve_model_files(domain, instance))
    hyper = [Hyperparameter("MODEL_WEIGHT_TUNE", -1.0, 4.0, power_10),
             Hyperparameter("LEARNING_RATE_TUNE", -5.0, 0.0, power_10)]
    kw = dict(plan="JaxStraightLinePlan", method_kwargs="{}", extra="", seconds=train_seconds)
    if method == "drp":
        kw.update(plan="JaxDeepReactivePolicy", method_kwargs="{'topology': [LAYER1_TUNE, LAYER1_TUNE]}")
        hyper.append(Hyperparameter("LAYER1_TUNE", 3.0, 8.0, power_2_int))
    if method == "replan":
        kw.update(extra="rollout_horizon=ROLLOUT_HORIZON_TUNE\n", seconds=min(train_seconds, 1.0))
        hyper.append(Hyperparameter("ROLLOUT_HORIZON_TUNE", 1.0, min(40.0, float(model.horizon)),
                                    lambda x: int(round(x))))
    tuner = JaxParameterTuning(model, _TUNE_CFG.format(**kw), hyper, online=(method == "replan"),
                               eval_trials=trials, gp_iters=iters, num_workers=workers)
    tuner.tune(key=42, log_file=f"gp_{method}_{os.path.basename(domain)}_{instance}.csv")
    if filepath:
        with open(filepath, "w") as f:
            f.write(tuner.best_config)
    print(tuner.best_config)
    return tuner


def main(argv=None):
    parser = argparse.ArgumentParser(prog="jaxplan", description="command line of the B200 JaxPlan planner")
    sub = parser.add_subparsers(dest="jaxplan", required=True)
    plan = sub.add_parser("plan", help="execute the planner on a specified RDDL problem")
    plan.add_argument("domain", type=str, help="RDDL domain file, or the name of a bundled fixture")
    plan.add_argument("instance", type=str, help="RDDL instance file, or an instance name / number")
    plan.add_argument("method", type=str, help="slp, drp (offline), replan (online) or a .cfg file")
    plan.add_argument("-e", "--episodes", type=int, default=1, help="number of evaluation episodes")
    plan.add_argument("--train-seconds", type=float, default=None, help="override the training budget")
    tune = sub.add_parser("tune", help="tune the planner's hyper-parameters on a specified RDDL problem")
    tune.add_argument("domain", type=str)
    tune.add_argument("instance", type=str)
    tune.add_argument("method", type=str, help="slp, drp (offline) or replan (online)")
    tune.add_argument("-t", "--trials", type=int, default=5, help="evaluation trials per choice")
    tune.add_argument("-i", "--iters", type=int, default=20, help="iterations of Bayesian optimisation")
    tune.add_argument("-w", "--workers", type=int, default=4, help="suggestions evaluated per iteration")
    tune.add_argument("-f", "--filepath", type=str, default="", help="where to save the best config")
    tune.add_argument("--train-seconds", type=float, default=30.0, help="training budget of a trial")
    args = parser.parse_args(argv)
    if args.jaxplan == "plan":
        run_plan(args.domain, args.instance, args.method, args.episodes, args.train_seconds)
    else:
        run_tune(args.domain, args.instance, args.method, args.trials, args.iters, args.workers,
                 args.filepath, args.train_seconds)


if __name__ == "__main__":
    main()
