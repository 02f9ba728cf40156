"""Parity harness shared by the CPU (emulator) and GPU (CUDA kernel) tests.

A `backend` runs assembled programs: 'emu' = oracle/isa_emulator.py on numpy (checks the
lowering / differentiation / assembler on CPU), 'cuda' = the real kernels through the C ABI.
Both are compared with the torch-CPU oracle on the same seeded inputs and supplied noise.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import planner as OP
from oracle.isa_emulator import Emulator
from oracle.model_eval import OracleModel
from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
from pyrddlgym_jax_b200.core.layout import draw_noise, noise_as_list, unpack_fluent_major
from pyrddlgym_jax_b200.core.program import PlanSpec

from .helpers import fixture_model, init_params, initial_state, pack_params


def make_params(model, key, T, seed=42, scale=0.5):
    params = init_params(model, T, seed=seed, scale=scale)
    if key == "quadcopter":
        params["thrust"] = params["thrust"] + 9.81
    if key == "reservoir":
        params["release"] = params["release"].abs() * 10 + 5.
    if key == "powergen":
        params["curProd"] = params["curProd"] * 2 + 3.
    if key == "marsrover":          # power within the box, rovers near minerals harvest
        params["power-x"] = params["power-x"] * 0.2
        params["power-y"] = params["power-y"] * 0.2
        params["harvest"] = params["harvest"] * 4.0
    return params


def run_forward(backend, prog, B, T, s0, params_flat, noise, disc=None, mparams=None):
    """-> dict(reward [T,B], returns [B], err [T,B], tape, final) as numpy arrays."""
    S = prog.S
    if backend == "emu":
        bufs = {"S0": s0.numpy().copy(), "PARAMS": params_flat.numpy().copy(),
                "NOISE": noise.numpy().copy(), "REWARD": np.zeros(T * B, np.float32),
                "RETURNS": np.zeros(B, np.float32), "ERR": np.zeros(T * B, np.uint32),
                "TAPE": np.zeros(T * prog.tape_words * B, np.float32),
                "FINAL": np.zeros(S * B, np.float32),
                "DISC": None if disc is None else disc.numpy().copy()}
        Emulator(prog.blob).run(B, T, bufs, None if mparams is None else mparams.numpy())
        return {"reward": bufs["REWARD"].reshape(T, B), "returns": bufs["RETURNS"],
                "err": bufs["ERR"].reshape(T, B), "tape": bufs["TAPE"], "final": bufs["FINAL"]}
    from pyrddlgym_jax_b200.core.engine import DeviceProgram
    dev = torch.device("cuda:0")
    dp = DeviceProgram(prog, specialize=(backend == "jit"))
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
    out = {"reward": z(T * B), "returns": z(B), "err": z(T * B, dt=torch.int32),
           "tape": z(T * prog.tape_words * B), "final": z(S * B)}
    dp.forward(B, T, s0.to(dev), policy=params_flat.to(dev), noise=noise.to(dev).reshape(-1),
               mparams=None if mparams is None else mparams.to(dev),
               disc=None if disc is None else disc.to(dev),
               reward=out["reward"], returns=out["returns"], err=out["err"],
               tape=out["tape"], final=out["final"])
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res["reward"] = res["reward"].reshape(T, B)
    res["err"] = res["err"].view(np.uint32).reshape(T, B)
    return res


def run_backward(backend, prog, B, T, tape, params_flat, noise, gret, disc=None, mparams=None):
    """-> gradient [T, A] (sum of the per-warp partials)."""
    A = prog.A
    if backend == "emu":
        em = Emulator(prog.blob)
        nW = em.num_warps(B)
        bufs = {"TAPE": np.asarray(tape), "PARAMS": params_flat.numpy().copy(),
                "NOISE": noise.numpy().copy(), "GRET": gret.numpy().copy(),
                "GRADP": np.zeros(nW * T * A, np.float32),
                "DISC": None if disc is None else disc.numpy().copy()}
        em.run(B, T, bufs, None if mparams is None else mparams.numpy())
        return bufs["GRADP"].reshape(nW, T * A).sum(0).reshape(T, A)
    from pyrddlgym_jax_b200.core.engine import DeviceProgram, reduce_partials
    dev = torch.device("cuda:0")
    dp = DeviceProgram(prog, specialize=(backend == "jit"))
    nW = dp.num_warps(B)
    gradp = torch.zeros(nW * T * A, device=dev)
    grad = torch.zeros(T * A, device=dev)
    dp.backward(B, T, torch.as_tensor(tape).to(dev), gret.to(dev), policy=params_flat.to(dev),
                noise=noise.to(dev).reshape(-1),
                mparams=None if mparams is None else mparams.to(dev),
                disc=None if disc is None else disc.to(dev), gradp=gradp)
    reduce_partials(gradp, nW, T * A, grad)
    torch.cuda.synchronize()
    return grad.cpu().numpy().reshape(T, A)


def forward_case(backend, key, B, T, relaxed, gamma=1.0, seed=1, compiler_cls=None,
                 param_scale=0.5, given=None, wrap_non_bool=False, stochastic=False,
                 pooled_softmax=False, **ckw):
    """Run program + oracle on the same inputs; returns everything a test wants to compare."""
    m = fixture_model(key)
    if relaxed:
        comp = (compiler_cls or logic.DefaultJaxRDDLCompilerWithGrad)(m, **ckw)
    else:
        comp = JaxRDDLCompiler(m)
    noop_true = tuple(v for v in m.action_fluents if m.variable_ranges[v] == "bool"
                      and bool(np.ravel(np.asarray(m.action_fluents[v]))[0]))
    spec = PlanSpec(bounds=m.bounds, wrap_non_bool=wrap_non_bool, stochastic=stochastic,
                    pooled_softmax=pooled_softmax, softmax_weight=2.0, noop_true=noop_true)
    builder = comp.builder(spec, train_policy=relaxed, batch_hint=B)
    # the adjoint program is assembled first: it decides what the forward program tapes for it
    bwd = builder.backward(gamma=gamma) if relaxed else None
    fwd = builder.forward(write_tape=relaxed, gamma=gamma)
    params = make_params(m, key, T, scale=param_scale)
    if stochastic:      # log_sigma rows follow the action rows (core/program.py: param_layout)
        gls = torch.Generator().manual_seed(seed + 1000)
        for name in m.action_fluents:
            if m.variable_ranges[name] != "bool":
                params[name + "/log_sigma"] = torch.randn(
                    T, m.variable_size(name), generator=gls) * 0.3 - 1.5
    noise = draw_noise(fwd.noise_layout, T, B, torch.Generator().manual_seed(seed))
    if given is not None:        # committed golden inputs (tests/golden/*.npz)
        params = {k: torch.from_numpy(np.array(given["param_" + k])) for k in params}
        noise = torch.from_numpy(np.array(given["noise"])).reshape(noise.shape)
    s0 = initial_state(m, fwd.state_layout, B)
    disc = None if gamma == 1.0 else torch.pow(torch.tensor(gamma), torch.arange(T)).float()
    got = run_forward(backend, fwd, B, T, s0, pack_params(m, params), noise, disc)
    om = OracleModel(m, comp.relax_config())
    if relaxed:
        P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        if stochastic:
            def pol(t, fls, pnoise):
                return OP.slp_train_actions(m, P, t, wrap_non_bool=wrap_non_bool, bounds=m.bounds,
                                            stochastic=True, pnoise=pnoise,
                                            pooled_softmax=pooled_softmax, softmax_weight=2.0)
            pol.n_noise = len(OP.slp_policy_noise_spec(m, pooled_softmax=pooled_softmax))
        else:
            pol = lambda t, fls: OP.slp_train_actions(m, P, t, wrap_non_bool=wrap_non_bool,   # noqa: E731
                                                      bounds=m.bounds, pooled_softmax=pooled_softmax,
                                                      softmax_weight=2.0)
    else:
        P = params
        pol = lambda t, fls: OP.slp_test_actions(m, P, t, m.bounds,    # noqa: E731
                                                 wrap_non_bool=wrap_non_bool,
                                                 pooled_softmax=pooled_softmax, softmax_weight=2.0)
    noise_steps = [noise_as_list(fwd.noise_layout, noise, t, B) for t in range(T)]
    ref = OP.rollout(om, pol, B, T, noise_steps)
    spec_or = [(x[0], tuple(x[1])) for x in (ref["noise_spec"] or [])]
    if relaxed and stochastic:
        spec_or = OP.slp_policy_noise_spec(m, pooled_softmax=pooled_softmax) + spec_or
    spec_pr = [(k, tuple(s)) for k, s, _, _, _ in fwd.noise_layout]
    assert T == 0 or spec_or == spec_pr, "oracle and lowering disagree on the draw sites"
    final = unpack_fluent_major(fwd.state_layout, torch.from_numpy(got["final"].copy()), B)
    return dict(model=m, comp=comp, builder=builder, fwd=fwd, bwd=bwd, params=params, P=P, noise=noise,
                s0=s0, disc=disc, got=got, ref=ref, final=final, gamma=gamma)


def gradient_case(backend, key, B, T, gamma=1.0, compiler_cls=None, param_scale=0.5, given=None,
                  wrap_non_bool=False, stochastic=False, pooled_softmax=False, **ckw):
    c = forward_case(backend, key, B, T, True, gamma, compiler_cls=compiler_cls,
                     param_scale=param_scale, given=given, wrap_non_bool=wrap_non_bool,
                     stochastic=stochastic, pooled_softmax=pooled_softmax, **ckw)
    m = c["model"]
    bwd = c["bwd"]
    gret = torch.full((B,), -1.0 / B)
    grad = run_backward(backend, bwd, B, T, c["got"]["tape"], pack_params(m, c["params"]),
                        c["noise"], gret, c["disc"])
    loss = OP.mean_loss(c["ref"]["reward"], gamma)
    loss.backward()
    grad_ref = pack_params(m, {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                               for k, v in c["P"].items()}).numpy()
    c.update(bwd=bwd, grad=grad, grad_ref=grad_ref, loss_ref=float(loss))
    return c


# Tolerances are ELEMENT-WISE: |got - ref| <= rtol * |ref| + floor, with rtol = 1e-5 for states,
# rewards and returns and 1e-4 for gradients (BASELINE.json north_star) and an absolute floor
# floor_frac * rtol * max|ref| for the entries that are sums with cancellation: 1e-6 of the largest
# entry for states / rewards / returns (FLOOR_FWD: a Reservoir reward of -5.4 is the fp32 sum of
# penalties of size 1000, whose own last-bit differences are 6e-5) and 1e-6 of the largest entry for
# gradients (FLOOR_FRAC).
FLOOR_FRAC = 1e-2
FLOOR_FWD = 1e-1


def assert_close(got, ref, rtol, what, floor_frac=FLOOR_FRAC, scale_min=0.0):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    both_nan = np.isnan(got) & np.isnan(ref)
    same_inf = np.isinf(ref) & (got == ref)
    floor = floor_frac * rtol * max(scale_min, float(np.nanmax(np.abs(np.where(np.isfinite(ref), ref, 0.0)))))
    with np.errstate(invalid="ignore"):
        bad = ~(np.abs(got - ref) <= rtol * np.abs(ref) + floor) & ~both_nan & ~same_inf
    if bad.any():
        idx = np.argwhere(bad)
        err = np.where(bad, np.abs(got - ref) / (np.abs(ref) + floor), 0.0)
        w = np.unravel_index(np.argmax(err), err.shape)
        raise AssertionError(
            f"{what}: {int(bad.sum())} of {ref.size} entries outside rtol {rtol:g} + floor {floor:.3g}; "
            f"worst at {tuple(int(i) for i in w)}: got {got[w]!r} ref {ref[w]!r} "
            f"(|diff| {abs(got[w] - ref[w]):.3g}); first {idx[:5].tolist()}")


def check_forward(c, rtol=1e-5):
    """states / rewards / returns element-wise within rtol 1e-5 (fp32), bool / int fluents and the
    error words bit-exact."""
    ref, got = c["ref"], c["got"]
    r_ref = ref["reward"].detach().numpy().T
    assert_close(got["reward"], r_ref, rtol, "reward", FLOOR_FWD, scale_min=1.0)
    ret_ref = OP.returns_from_rewards(ref["reward"].detach(), c["gamma"]).numpy()
    assert_close(got["returns"], ret_ref, rtol, "returns", FLOOR_FWD, scale_min=float(np.abs(r_ref).max()) if r_ref.size else 0.0)
    for s in c["model"].state_fluents:
        a = ref["final"][s].detach().reshape(c["final"][s].shape[0], -1)
        b = c["final"][s].reshape(a.shape)
        if a.dtype in (torch.bool, torch.int32):
            assert torch.equal(a.to(torch.int32), b.to(torch.int32)), f"fluent {s} not bit-exact"
        else:
            assert_close(b.numpy(), a.numpy(), rtol, f"final state {s}", FLOOR_FWD, scale_min=1.0)
    np.testing.assert_array_equal(got["err"], ref["error"].numpy().T.astype(np.uint32))


def check_gradient(c, rtol=1e-4):
    """plan gradient element-wise within rtol 1e-4 (fp32)."""
    assert_close(c["grad"], c["grad_ref"], rtol, "gradient")
