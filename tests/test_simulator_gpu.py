"""JaxRDDLSimulator (core/simulator.py:43-254 of the reference: the gym simulation backend) on the exact
rollout kernels: one launch per environment step, constraints evaluated inside the same step graph."""
import numpy as np
import pytest
import torch

from oracle.model_eval import OracleModel, NoiseCursor
from pyrddlgym_jax_b200.rddl.model import load_model

from .helpers import fixture_model

pytestmark = pytest.mark.gpu

COUNTER_RDDL = """
domain counter_toy {
    types { cell : object; };
    pvariables {
        STEP(cell) : { non-fluent, real, default = 1.0 };
        x(cell)    : { state-fluent, real, default = 0.0 };
        hits       : { state-fluent, int, default = 0 };
        big(cell)  : { interm-fluent, bool };
        u(cell)    : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        big(?c) = x(?c) + STEP(?c) * u(?c) > 1.5;
        x'(?c)  = x(?c) + STEP(?c) * u(?c);
        hits'   = hits + (sum_{?c : cell} [ big(?c) ]);
    };
    reward = - (sum_{?c : cell} [ pow[x'(?c) - 2.0, 2] ]);
    state-invariants {
        forall_{?c : cell} [ x(?c) >= -0.5 ];
    };
    action-preconditions {
        forall_{?c : cell} [ u(?c) <= 1.0 ];
        forall_{?c : cell} [ u(?c) >= -1.0 ];
    };
    termination {
        exists_{?c : cell} [ x(?c) > 2.5 ];
    };
}
non-fluents nf { domain = counter_toy; objects { cell : {c1, c2}; }; non-fluents { STEP(c2) = 0.5; }; }
instance inst { domain = counter_toy; non-fluents = nf; init-state { x(c1) = 0.25; };
    max-nondef-actions = pos-inf; horizon = 10; discount = 1.0; }
"""


def test_simulator_steps_constraints_and_termination(cuda_device):
    from pyrddlgym_jax_b200.core.simulator import (JaxRDDLSimulator,
                                                   RDDLActionPreconditionNotSatisfiedError,
                                                   RDDLStateInvariantNotSatisfiedError)
    m = load_model(COUNTER_RDDL)
    sim = JaxRDDLSimulator(m, key=3)
    obs, done = sim.reset()
    assert obs == {"x___c1": 0.25, "x___c2": 0.0, "hits": 0} and not done
    assert sim.check_state_invariants()
    assert sim.check_action_preconditions({"u___c1": 1.0})
    with pytest.raises(RDDLActionPreconditionNotSatisfiedError):
        sim.check_action_preconditions({"u___c1": 1.5})
    assert not sim.check_action_preconditions({"u": np.array([0.0, -2.0])}, silent=True)
    x, hits, total = np.array([0.25, 0.0]), 0, 0.0
    for t in range(4):
        u = np.array([1.0, 0.8])
        obs, reward, done = sim.step({"u___c1": 1.0, "u___c2": 0.8})
        hits += int(np.sum(x + np.array([1.0, 0.5]) * u > 1.5))
        x = x + np.array([1.0, 0.5]) * u
        assert abs(obs["x___c1"] - x[0]) < 1e-6 and abs(obs["x___c2"] - x[1]) < 1e-6
        assert obs["hits"] == hits
        assert abs(reward - (-np.sum((x - 2.0) ** 2))) < 1e-5
        assert done == bool(np.any(x > 2.5))
        assert sim.check_state_invariants()
        if done:
            break
    assert done and t == 2
    obs, done = sim.reset()
    assert obs["x___c1"] == 0.25 and not done
    sim.step({"u": np.array([-1.0, 0.0])})              # x(c1) = -0.75 breaks the invariant
    with pytest.raises(RDDLStateInvariantNotSatisfiedError):
        sim.check_state_invariants()
    assert not sim.check_state_invariants(silent=True)
    # tensors, several environments in one launch
    vec = JaxRDDLSimulator(m, key=3, keep_tensors=True, n_envs=5)
    obs, _ = vec.reset()
    assert obs["x"].shape == (5, 2)
    obs, reward, done = vec.step({"u": np.array([0.5, 0.5])})
    np.testing.assert_allclose(obs["x"], np.tile([0.75, 0.25], (5, 1)), atol=1e-6)
    assert reward.shape == (5,)


@pytest.mark.parametrize("key", ["quadcopter", "reservoir", "pomdp"])
def test_simulator_matches_the_oracle_step(cuda_device, key):
    """Deterministic transitions equal the oracle's exact step; stochastic ones are reproducible per seed,
    differ across seeds and keep the reward / state finite; a POMDP hands out its observ-fluents."""
    from pyrddlgym_jax_b200.core.simulator import JaxRDDLSimulator
    m = fixture_model(key)
    sim = JaxRDDLSimulator(m, key=11, keep_tensors=True)
    obs, _ = sim.reset()
    g = np.random.default_rng(0)
    acts = {a: (np.asarray(m.init_values[a]) + 0.1 * g.random(np.asarray(m.init_values[a]).shape))
            .astype(np.float32) for a in m.action_fluents
            if m.variable_ranges[a] == "real"}
    traj = []
    for _ in range(3):
        obs, reward, done = sim.step(acts)
        traj.append((obs, reward))
        assert np.isfinite(reward)
    if key == "pomdp":
        assert set(obs) == {"sensor", "alarm"} and obs["alarm"].dtype == bool
        assert set(sim.state) == {"level"}
    if key == "quadcopter":
        om = OracleModel(m, None)
        fls, nfls = om.init_fls_nfls(1)
        a_t = {k: torch.as_tensor(v).unsqueeze(0) for k, v in acts.items()}
        for t in range(3):
            for a in m.action_fluents:
                a_t.setdefault(a, torch.as_tensor(np.asarray(m.init_values[a])).unsqueeze(0))
            fls, reward, _ = om.step(fls, nfls, a_t, None, 1)[:3]
            for s in m.state_fluents:
                np.testing.assert_allclose(traj[t][0][s], fls[s][0].numpy(), rtol=1e-5, atol=1e-6)
            assert abs(traj[t][1] - float(reward[0])) <= 1e-5 * max(1.0, abs(float(reward[0])))
    else:
        again = JaxRDDLSimulator(m, key=11, keep_tensors=True)
        again.reset()
        other = JaxRDDLSimulator(m, key=12, keep_tensors=True)
        other.reset()
        o2 = o3 = None
        for _ in range(3):
            o2, r2, _ = again.step(acts)
            o3, r3, _ = other.step(acts)
        for k_ in obs:
            np.testing.assert_array_equal(np.asarray(obs[k_]), np.asarray(o2[k_]))
        assert any(not np.array_equal(np.asarray(obs[k_]), np.asarray(o3[k_])) for k_ in obs)


def test_controller_in_the_loop(cuda_device):
    """An offline controller acting in the simulator for an episode: the loop a user of the reference runs
    through pyRDDLGym's environment (examples/run_plan.py:57-61)."""
    from pyrddlgym_jax_b200.core.planner import (JaxBackpropPlanner, JaxOfflineController,
                                                 JaxStraightLinePlan)
    from pyrddlgym_jax_b200.core.simulator import JaxRDDLSimulator
    m = fixture_model("reservoir")
    planner = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=64, rollout_horizon=8,
                                 optimizer_kwargs={"learning_rate": 0.1}, pgpe=None)
    ctrl = JaxOfflineController(planner, key=5, train_on_reset=True, epochs=30, print_summary=False,
                                print_progress=False)
    sim = JaxRDDLSimulator(m, key=7, keep_tensors=True)
    state, done = sim.reset()
    ctrl.reset()
    total = 0.0
    for _ in range(8):
        action = ctrl.sample_action(state)
        assert sim.check_action_preconditions(action, silent=True)
        state, reward, done = sim.step(action)
        total += reward
    assert np.isfinite(total)


def test_tuning_on_a_small_budget(cuda_device, tmp_path):
    """JaxParameterTuning with the real objective (train, then act in the simulator): a few iterations on
    Reservoir with a tiny budget; the best configuration builds a controller that acts."""
    from pyrddlgym_jax_b200.core.tuning import Hyperparameter, JaxParameterTuning, power_10
    template = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
sigmoid_weight=MODEL_WEIGHT_TUNE

[Planner]
method='JaxStraightLinePlan'
method_kwargs={}
optimizer='rmsprop'
optimizer_kwargs={'learning_rate': LEARNING_RATE_TUNE}
batch_size_train=32
batch_size_test=32
rollout_horizon=8
pgpe=None

[Optimize]
epochs=25
train_seconds=5
print_summary=False
print_progress=False
"""
    m = fixture_model("reservoir")
    tuner = JaxParameterTuning(m, template, [Hyperparameter("MODEL_WEIGHT_TUNE", 0.0, 3.0, power_10),
                                             Hyperparameter("LEARNING_RATE_TUNE", -3.0, 0.0, power_10)],
                               online=False, eval_trials=1, rollouts_per_trial=1, gp_iters=4, verbose=False)
    best = tuner.tune(key=1, log_file=str(tmp_path / "log.csv"))
    assert set(best) == {"MODEL_WEIGHT_TUNE", "LEARNING_RATE_TUNE"} and np.isfinite(tuner.best_target)
    assert len(tuner.log) == 4
    agent = tuner.build_best_controller()
    from pyrddlgym_jax_b200.core.simulator import JaxRDDLSimulator
    sim = JaxRDDLSimulator(m, key=2, keep_tensors=True)
    state, _ = sim.reset()
    agent.reset()
    state, reward, _ = sim.step(agent.sample_action(state))
    assert np.isfinite(reward)
