"""GPU parity tests: the CUDA interpreter kernels, called through the C ABI, against the
torch-CPU oracle on the same seeded inputs and supplied noise.

Tolerances (BASELINE.json north_star): bool / int fluents bit-exact in exact mode; float
states and returns rtol 1e-5; gradients rtol 1e-4 (fp32)."""
import numpy as np
import pytest
import torch

from pyrddlgym_jax_b200.core import logic

from .parity import check_forward, check_gradient, forward_case, gradient_case

pytestmark = pytest.mark.gpu
KEYS = ["quadcopter", "reservoir", "wildfire", "marsrover", "powergen", "discrete",
        "distributions", "ops", "samplers", "matrix"]
# 'cuda' = generic interpreter kernel, 'jit' = per-program specialised kernel (core/codegen.py)
BACKENDS = ["cuda", "jit"]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("relaxed", [False, True])
def test_forward_cuda(cuda_device, backend, key, relaxed):
    check_forward(forward_case(backend, key, B=37, T=8, relaxed=relaxed))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("gamma", [1.0, 0.9])
def test_gradient_cuda(cuda_device, backend, key, gamma):
    check_gradient(gradient_case(backend, key, B=29, T=6, gamma=gamma))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 257])
def test_ragged_batches_cuda(cuda_device, backend, B):
    """Batch sizes around the rollouts-per-warp / warps-per-CTA boundaries."""
    check_gradient(gradient_case(backend, "reservoir", B=B, T=3))
    check_forward(forward_case(backend, "wildfire", B=B, T=3, relaxed=False))


@pytest.mark.parametrize("key,B", [("quadcopter", 4096), ("wildfire", 2048), ("reservoir", 32)])
def test_large_batch_lane_layouts_jit(cuda_device, key, B):
    """Batch sizes that make the assembler pick other lanes-per-rollout layouts."""
    c = gradient_case("jit", key, B=B, T=4)
    check_forward(c)
    check_gradient(c)


def test_empty_inputs_cuda(cuda_device):
    """B = 0 and T = 0 are accepted and touch nothing."""
    c = forward_case("cuda", "reservoir", B=5, T=0, relaxed=True)
    assert np.all(c["got"]["returns"] == 0)
    from pyrddlgym_jax_b200.core.engine import DeviceProgram
    dp = DeviceProgram(c["fwd"])
    dp.forward(0, 4, torch.zeros(1, device="cuda"))
    torch.cuda.synchronize()


@pytest.mark.parametrize("backend", BACKENDS)
def test_horizon_full_cuda(cuda_device, backend):
    """Full-horizon Quadcopter (T=100) against the oracle, relaxed model + gradient, from the
    planner's own initial plan (normal(0.01), planner.py:636).  (With O(1) random torques the
    100-step attitude dynamics are chaotic and fp32 round-off differences between any two
    libm implementations are amplified beyond 1e-4; that regime is covered at T <= 8.)"""
    c = gradient_case(backend, "quadcopter", B=64, T=100, param_scale=0.01)
    check_forward(c)
    check_gradient(c)


class _Godel(logic.SigmoidRelational, logic.GodelNormLogical, logic.LinearIfElse,
             logic.GumbelSigmoidBernoulli, logic.SoftFloor, logic.SoftRound):
    pass


@pytest.mark.parametrize("backend", BACKENDS)
def test_variants_cuda(cuda_device, backend):
    c = gradient_case(backend, "wildfire", B=19, T=5, compiler_cls=_Godel,
                      sigmoid_weight=3.0, use_logic_ste=True, use_ste_bernoulli=True)
    check_forward(c)
    check_gradient(c)
    c = gradient_case(backend, "wildfire", B=19, T=5, sigmoid_weight=100.0,
                      bernoulli_sigmoid_weight=100.0)
    check_forward(c)
    check_gradient(c)


@pytest.mark.parametrize("key", KEYS + ["gammafluent"])
def test_thread_per_rollout_kernels_cuda(cuda_device, key, monkeypatch):
    """lanes = 1 programs (a lane owns a whole rollout, core/codegen_tpr.py): generated kernel and
    interpreter, exact forward + relaxed forward + gradient, ragged batch."""
    from pyrddlgym_jax_b200.core.program import ProgramError
    monkeypatch.setenv("RK_LANES", "1")
    for backend in BACKENDS:
        try:
            c = gradient_case(backend, key, B=77, T=5, gamma=0.95)
        except ProgramError:
            pytest.skip("the program does not fit a lanes = 1 layout")
        assert c["fwd"].lanes == 1 and c["bwd"].lanes == 1
        check_forward(c)
        check_gradient(c)
        check_forward(forward_case(backend, key, B=77, T=5, relaxed=False))


def test_thread_per_rollout_large_batch_cuda(cuda_device):
    """The assembler's own choice at a batch that fills the machine: MarsRover picks lanes = 1."""
    c = gradient_case("jit", "marsrover", B=40000, T=6)
    assert c["bwd"].lanes == 1 and c["fwd"].lanes == 1
    check_forward(c)
    check_gradient(c)


class _Luka(logic.SigmoidRelational, logic.LukasiewiczNormLogical, logic.LinearIfElse,
            logic.DeterminizedBernoulli):
    pass


@pytest.mark.parametrize("backend", BACKENDS)
def test_lukasiewicz_cuda(cuda_device, backend):
    """The Lukasiewicz t-norm (logic.py:674-778) on the Wildfire / Reservoir CPFs."""
    for key in ("wildfire", "reservoir"):
        c = gradient_case(backend, key, B=19, T=5, compiler_cls=_Luka, sigmoid_weight=2.0)
        check_forward(c)
        check_gradient(c)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("variant", range(6))
def test_ops_variants_cuda(cuda_device, backend, variant):
    """Every relaxation family with and without straight-through estimators, under the product,
    Goedel and Lukasiewicz t-norms, on the fixtures that use every operator and sampler of the
    instruction set (tests/test_lowering_cpu.py::OPS_VARIANTS)."""
    from .test_lowering_cpu import OPS_VARIANTS
    cls, kw = OPS_VARIANTS[variant]
    for key in ("ops", "samplers"):
        c = gradient_case(backend, key, B=23, T=5, compiler_cls=cls, **kw)
        check_forward(c)
        check_gradient(c)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("key", ["reservoir", "quadcopter"])
def test_wrap_non_bool_cuda(cuda_device, backend, key):
    """Box-wrapped real actions (planner.py:620-628) in the train and the test policy."""
    check_forward(forward_case(backend, key, B=21, T=5, relaxed=False, wrap_non_bool=True))
    c = gradient_case(backend, key, B=21, T=5, wrap_non_bool=True)
    check_forward(c)
    check_gradient(c)


def test_deterministic_cuda(cuda_device):
    """Two runs give bit-identical gradients (fixed reduction order, no float atomics)."""
    a = gradient_case("jit", "wildfire", B=300, T=5)["grad"]
    b = gradient_case("jit", "wildfire", B=300, T=5)["grad"]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("method", ["sogbofa", "sorting"])
@pytest.mark.parametrize("wrap_sigmoid", [True, False])
def test_concurrency_projection_cuda(cuda_device, method, wrap_sigmoid):
    """rk_project_slp against the oracle's restatement of planner.py:489-617, 947-954 on the
    Wildfire plan (50 boolean actions, max-nondef-actions = 1 and 3)."""
    from oracle import planner as OP
    from pyrddlgym_jax_b200.core import engine
    from .helpers import fixture_model, pack_params
    m = fixture_model("wildfire")
    T, w, pmin = 7, 2.5, 1e-3
    g = torch.Generator().manual_seed(5)
    for allowed in (1, 3):
        if wrap_sigmoid:
            params = {a: torch.randn(T, m.variable_size(a), generator=g) * 1.5
                      for a in m.action_fluents}
        else:
            params = {a: torch.rand(T, m.variable_size(a), generator=g) * 0.8 + 0.1
                      for a in m.action_fluents}
        params["put-out"][0] = -3.0 if wrap_sigmoid else 0.05        # a step with nothing active
        boxed = OP.project_box(m, params, m.bounds, wrap_sigmoid, pmin, w)
        want, _ = OP.project_max_constraint(m, boxed, method, allowed, wrap_sigmoid, pmin, w, 100)
        flat = pack_params(m, boxed).to(cuda_device).contiguous()
        segs, off = [], 0
        for a in m.action_fluents:
            n = m.variable_size(a)
            segs.append((off, n, False, w))
            off += n
        conv = torch.ones(1, dtype=torch.int32, device=cuda_device)
        engine.project_slp(method, T, off, flat.view(-1), segs, allowed, 100, wrap_sigmoid, pmin, conv)
        torch.cuda.synchronize()
        got = flat.cpu().numpy()
        ref = pack_params(m, want).numpy()
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)
        # the projected plan satisfies the constraint under the test policy
        thr = 0.0 if wrap_sigmoid else 0.5
        assert int((got > thr).sum(1).max()) <= allowed
