#!/usr/bin/env python
"""bench.py -- planner iterations of the BASELINE.json workloads on N B200s (one node).

One bench "step" = one planner iteration as the reference runs it (planner.py:3245-3260, pgpe=None):
relaxed forward rollouts writing the state tape, loss, backprop-through-time, gradient reduction
(+ one all-reduce when N > 1), rmsprop step + projections, then the exact-model test rollouts.

Headline (``value``, ``e2e``, ``roofline``): BASELINE.json configs[4], the largest configuration --
MarsRover instance 5, straight-line plan, a GLOBAL batch of 262144 rollouts sharded over the N GPUs
(``"scaling": "strong"``: N = 1 runs all 262144 rollouts on one GPU).  The same JSON line carries a
``workloads`` array with ``value``, ``e2e``, ``ms_per_step``, ``roofline`` and (N = 1) ``cpu_baseline``
of ALL FIVE configs: Reservoir (B 32), Quadcopter (B 4096), Wildfire (B 16384), PowerGen DRP 2x256
(B 32768) -- those four with their batch PER GPU (weak scaling) -- and the MarsRover headline itself.

``value`` = train rollout-steps per second (global B*T per iteration) with inputs resident in HBM;
``e2e`` = the same through the public planner API with the initial state and the plan copied from
pinned host memory and the losses + plan read back every step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

``--workload NAME`` runs only that workload (it becomes the line's headline; no ``workloads`` array).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = "rollout_steps_per_sec"
UNIT = "rollout-steps/s"
HEADLINE = "marsrover"

# name -> BASELINE.json configs[i]; hyper-parameters from the reference's shipped cfg of the domain.
# B = rollouts per GPU ("weak") or in total ("strong": split over the GPUs of the run).
WORKLOADS = {
    # configs[0] -- the reference's own CPU-sized case (launch-latency bound on a GPU)
    "reservoir": dict(
        label="Reservoir_Continuous instance 1, slp, batch 32, horizon 40, fp32, rmsprop lr 0.2, pgpe=None",
        domain="Reservoir_Continuous", instance="instance1", B=32, T=40, lr=0.2, scaling="weak",
        compiler="DefaultJaxRDDLCompilerWithGrad", compiler_kwargs={}, plan_kwargs={}),
    # configs[1]
    "quadcopter": dict(
        label="Quadcopter instance 1, slp, batch 4096/GPU, horizon 100, fp32, rmsprop lr 0.03, pgpe=None",
        domain="Quadcopter", instance="instance1", B=4096, T=100, lr=0.03, scaling="weak",
        compiler="DefaultJaxRDDLCompilerWithGrad", compiler_kwargs={"sigmoid_weight": 10.},
        plan_kwargs={}),
    # configs[2]
    "wildfire": dict(
        label="Wildfire_MDP_ippc2014 instance 5, slp, product t-norm + Gumbel-sigmoid Bernoulli "
              "relaxations (weights 100/100), batch 16384/GPU, horizon 40, rmsprop lr 0.01, "
              "SOGBOFA concurrency projection, pgpe=None",
        domain="Wildfire_MDP_ippc2014", instance="instance5", B=16384, T=40, lr=0.01, scaling="weak",
        compiler="GumbelBernoulliJaxRDDLCompilerWithGrad",
        compiler_kwargs={"sigmoid_weight": 100., "bernoulli_sigmoid_weight": 100.},
        plan_kwargs={"sigmoid_weight": 10.0}),
    # configs[3] -- deep reactive policy, 2x256 tanh MLP (the only dense contraction of the path)
    "powergen_drp": dict(
        label="PowerGen_Continuous instance 1 (5 plants), drp 2x256 tanh MLP (stochastic heads), "
              "batch 32768/GPU, horizon 100, fp32, rmsprop lr 1e-4, pgpe=None",
        domain="PowerGen_Continuous", instance="instance1", B=32768, T=100, lr=1e-4, scaling="weak",
        compiler="DefaultJaxRDDLCompilerWithGrad", compiler_kwargs={},
        plan="drp", plan_kwargs={"topology": [256, 256]}),
    # configs[4] -- large-object instance; a global batch of 262144 split over the GPUs of the run
    "marsrover": dict(
        label="MarsRover_ippc2023 instance 5 (4 rovers x 8 minerals), slp, GLOBAL batch 262144 "
              "sharded over the GPUs, horizon 60, fp32, rmsprop lr 0.01, pgpe=None",
        domain="MarsRover_ippc2023", instance="instance5", B=262144, T=60, lr=0.01, scaling="strong",
        compiler="DefaultJaxRDDLCompilerWithGrad", compiler_kwargs={"sigmoid_weight": 10.},
        plan_kwargs={}),
}
ORDER = ["reservoir", "quadcopter", "wildfire", "powergen_drp", "marsrover"]


def workload_model(name: str):
    from pyrddlgym_jax_b200.rddl.model import load_model
    w = WORKLOADS[name]
    d = os.path.join(ROOT, "pyrddlgym_jax_b200", "domains", w["domain"])
    return load_model(os.path.join(d, "domain.rddl"), os.path.join(d, w["instance"] + ".rddl"))


def per_gpu_batch(name: str, n_gpus: int) -> int:
    w = WORKLOADS[name]
    return w["B"] // n_gpus if w["scaling"] == "strong" else w["B"]


def workload_config(name: str, n_gpus: int) -> dict:
    """The ``config`` object of a line: identical for the B200 arm and the reference arm."""
    w = WORKLOADS[name]
    B = per_gpu_batch(name, n_gpus)
    return {"workload": w["label"], "global_batch": n_gpus * B, "batch_per_gpu": B,
            "horizon": w["T"], "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single"}


# ---------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, busy = [], 0.0, set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                util = float(r[7])
                mx = max(mx, float(r[1]))
                if util < 50.0:             # only samples taken under load count
                    continue
                busy += 1
                sm.append(float(r[0]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(self.rows), "samples_under_load": busy}


# ---------------------------------------------------------------------------------------
def make_compiler(name: str, model):
    from pyrddlgym_jax_b200.core import logic
    w = WORKLOADS[name]
    return getattr(logic, w["compiler"])(model, **w["compiler_kwargs"])


def oracle_iteration_fn(name, model, B, T):
    """One planner iteration of the CPU oracle (torch-CPU fp32 restatement of the reference):
    relaxed rollouts + loss + autograd BPTT + rmsprop + projections + exact test rollouts.
    Used by the CPU legs only (cpu_baseline, --impl reference)."""
    from oracle.iteration import OracleIteration
    from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
    from pyrddlgym_jax_b200.core.program import PlanSpec
    w = WORKLOADS[name]
    pw = float(w["plan_kwargs"].get("sigmoid_weight", 1.0))
    comp = make_compiler(name, model)
    spec = PlanSpec(bounds=model.bounds, sigmoid_weight=pw,
                    kind="external" if w.get("plan") == "drp" else "slp")
    lay_train = comp.builder(spec, True, (), B).forward(True).noise_layout
    lay_test = JaxRDDLCompiler(model).builder(spec, False, (), B).forward(False).noise_layout
    return OracleIteration(model, comp.relax_config(), (lay_train, lay_test), B, T, w["lr"],
                           plan="drp" if w.get("plan") == "drp" else "slp", plan_sigmoid_weight=pw,
                           topology=w["plan_kwargs"].get("topology", ()), ent_coeff=0.1)


# measured in round 1 on the GPU box's host (16-24 threads): rollout-steps/s of the CPU port; used
# only to size the bounded CPU sample so that one CPU step takes about a second
CPU_RATE_HINT = {"reservoir": 1.1e4, "quadcopter": 1.8e5, "wildfire": 1.9e5, "powergen_drp": 4.0e5,
                 "marsrover": 3.2e5}


def cpu_sample_batch(name: str, n_gpus: int = 1) -> int:
    """Rollouts per CPU step: a bounded sample of the workload's (global) batch sized for about one
    second per step, so that K + W steps of every workload end within minutes."""
    w = WORKLOADS[name]
    B_full = per_gpu_batch(name, n_gpus) * n_gpus
    b = int(CPU_RATE_HINT[name] * 1.0 / w["T"])
    b = max(32, 1 << int(np.floor(np.log2(max(b, 32)))))
    return min(B_full, b)


def cpu_leg(name: str, n_steps: int, n_warm: int, budget_s: float = 0.0, n_gpus: int = 1) -> dict:
    """Times the CPU oracle's planner iteration on all host threads (kind 'port': the reference itself
    -- jax + pyRDDLGym + rddlrepository -- cannot be installed offline)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = workload_model(name)
    B, T = cpu_sample_batch(name, n_gpus), WORKLOADS[name]["T"]
    it = oracle_iteration_fn(name, model, B, T)
    for _ in range(n_warm):
        it()
    n, t0 = 0, time.perf_counter()
    while n < n_steps and (budget_s <= 0 or n == 0 or time.perf_counter() - t0 < budget_s):
        it()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": B * T / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "steps": n, "warmup": n_warm, "ms_per_step": dt * 1e3,
            "sample": f"{n} planner iterations of {B} rollouts x {T} steps (a bounded sample of the "
                      f"workload's {per_gpu_batch(name, n_gpus) * n_gpus}) on the torch-CPU oracle "
                      f"restatement, {cores} threads"}


def run_reference(args):
    """--impl reference: the reference's planner iteration on the host cores, same config / metric /
    unit / steps / warm-up as the B200 arm; every step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or HEADLINE
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    W = max(args.warmup, 3)
    leg = cpu_leg(name, args.steps, W, n_gpus=n_gpus)
    # XLA's CPU code for these element-wise graphs is essentially single-threaded, so
    # a 1-thread figure is reported beside the all-threads one
    cores = leg["cores"]
    torch.set_num_threads(1)
    one = cpu_leg(name, 2, 1, n_gpus=n_gpus)
    torch.set_num_threads(cores)
    leg["cores"] = cores
    leg["value_one_thread"] = one["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT,
        "n_gpus": n_gpus, "steps": args.steps, "warmup": W, "ms_per_step": leg["ms_per_step"],
        "higher_is_better": True, "scaling": WORKLOADS[name]["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(name, n_gpus),
        "note": "CPU restatement (oracle/), not jax: jax, pyRDDLGym and rddlrepository are not "
                "installable offline; each step is a bounded sample of the workload (cpu_baseline.sample)",
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample",
                                             "value_one_thread")},
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# figures read from the committed ncu captures of the same workload (profiles/), per launch of the
# adjoint kernel at the per-GPU batch of the N = 1 run; None = not captured at this size
PROFILED = {}
try:
    with open(os.path.join(ROOT, "profiles", "r02_profiled.json")) as _f:
        PROFILED = json.load(_f)
except Exception:
    pass


class Ctx:
    pass


def measure_workload(name: str, ctx: Ctx, steps: int, warmup: int, burst_s: float = 0.0,
                     with_cpu: bool = False) -> dict:
    """Builds the planner of one workload and measures it; returns the fields of a JSON line."""
    from pyrddlgym_jax_b200.core.planner import (JaxBackpropPlanner, JaxStraightLinePlan,
                                                 JaxDeepReactivePolicy)
    from pyrddlgym_jax_b200.core import engine, logic
    dev, world, n_gpus = ctx.dev, ctx.world, ctx.world
    wl = WORKLOADS[name]
    model = workload_model(name)
    B, T = (ctx.batch or per_gpu_batch(name, n_gpus)), wl["T"]
    plan_cls = JaxDeepReactivePolicy if wl.get("plan") == "drp" else JaxStraightLinePlan
    planner = JaxBackpropPlanner(
        model, plan_cls(**wl["plan_kwargs"]), batch_size_train=B, batch_size_test=B,
        rollout_horizon=T, optimizer="rmsprop", optimizer_kwargs={"learning_rate": wl["lr"]},
        pgpe=None, compiler=getattr(logic, wl["compiler"]), compiler_kwargs=wl["compiler_kwargs"],
        device=dev)
    ctx.keep.append(planner)        # stays alive: its CUDA graph may hold a collective
    planner.initialize(42)
    fw, bw, te = planner.train_fwd.program, planner.train_bwd.program, planner.test_fwd.program
    S, A, N = fw.S, fw.A, fw.N
    has_noise = planner.train_noise is not None or planner.test_noise is not None \
        or (planner.drp is not None and planner.drp.pol_noise is not None)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        # supplied-noise mode: fresh draws every iteration, like the reference's key splits
        if has_noise:
            planner.redraw_noise()
        planner.iterate()

    # ---- warm-up (also captures the CUDA graph) ------------------------------------------
    W = max(warmup, 3)
    for _ in range(W):
        step()
    barrier()
    if burst_s > 0:       # untimed burst so that the clock sampler sees the GPU under this load
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < burst_s:
            for _ in range(8):
                step()
            torch.cuda.synchronize()
        barrier()

    # ---- device-resident timed loop: EXACTLY `steps` iterations ------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    t_local = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_local.item()) / steps
    value = n_gpus * B * T / (ms_per_step * 1e-3)

    # ---- e2e: public API with host buffers, H2D + D2H inside the timed region --------------
    # inputs of one optimize() call as the public API takes them: the initial state (`subs`,
    # un-batched: S words per model, tiled to the B rollouts on the device) and the plan to start
    # from (`guess`); outputs: the three status scalars and the plan (callback['params']).  The host
    # copies are nodes of the same CUDA graph as the iteration (JaxBackpropPlanner.iterate_from_host)
    init_train_h, init_test_h = planner.init_train_host, planner.init_test_host
    params_h = planner.params.cpu().pin_memory()
    h2d = (init_train_h.numel() + init_test_h.numel() + params_h.numel()) * 4
    d2h = planner._stats_host.numel() * 4 + params_h.numel() * 4
    for _ in range(3):
        planner.iterate_from_host(init_train_h, init_test_h, params_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        if has_noise:
            planner.redraw_noise()
        _, _, _, pback_h = planner.iterate_from_host(init_train_h, init_test_h, params_h)
        params_h.copy_(pback_h)
    t_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * B * T / (float(t_e2e.item()) / steps * 1e-3)

    # ---- per-kernel timing (CUDA events on the launch stream, L2 flushed between launches) --
    def time_kernel(fn, reps=10):
        ts = []
        for _ in range(reps):
            ctx.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    hbm_peak, peak_src = ctx.hbm_peak, ctx.peak_src
    prof = PROFILED.get(name, {})
    if planner.is_drp:
        # closed-loop plan: the iteration is 2T decision epochs of a few launches each; time the
        # sweeps and the dominant kernel (the hidden-layer GEMM, the only dense contraction)
        drp = planner.drp
        t_upd = time_kernel(lambda: drp.update_kernels(0), reps=3)
        t_test = time_kernel(lambda: drp.test_kernels(0), reps=3)
        roofline, kernels_ms, n_launch = drp.bench_roofline(time_kernel, ctx.peaks, peak_src)
        kernels_ms.update({"train_update_sweeps": t_upd, "exact_test_sweep": t_test})
        roofline_other = {}
        test_steps_per_s = B * T / (t_test * 1e-3)
        bytes_note = f"activation tape {drp.tape_bytes() / 1e9:.2f} GB"
    else:
        p0 = planner.params[0]

        def k_fwd():
            planner.train_fwd.forward(B, T, planner.s0_train, policy=p0, noise=planner.train_noise,
                                      mparams=planner.train_mparams, disc=planner.disc,
                                      reward=planner.train_reward[0], returns=planner.train_returns[0],
                                      err=planner.train_err[0], tape=planner.tape)

        def k_bwd():
            planner.train_bwd.backward(B, T, planner.tape, planner.gret, policy=p0,
                                       noise=planner.train_noise, mparams=planner.train_mparams,
                                       disc=planner.disc, gradp=planner.gradp)

        def k_test():
            planner.test_fwd.forward(B, T, planner.s0_test, policy=p0, noise=planner.test_noise,
                                     disc=planner.disc, reward=planner.test_reward[0],
                                     returns=planner.test_returns[0], err=planner.test_err[0],
                                     final=planner.test_final[0])

        def k_red():
            engine.reduce_partials(planner.gradp, planner.nwarps, T * A, planner.grad[0])

        t_fwd, t_bwd, t_test, t_red = (time_kernel(k) for k in (k_fwd, k_bwd, k_test, k_red))
        # algorithmic bytes per launch: adjoint = tape + noise read once per
        # rollout-step + dL/dreturn + the gradient partial rows it writes
        bytes_bwd = B * T * 4 * (fw.tape_words + N) + 4 * B + planner.nwarps * T * A * 4
        bytes_fwd = B * T * (4 * fw.tape_words + 4 * N + 4 + 4) + 4 * S * B + 4 * B
        bytes_test = B * T * (4 * te.N + 4 + 4) + 4 * te.S * B * 2 + 4 * B
        achieved = bytes_bwd / (t_bwd * 1e-3) / 1e9
        test_steps_per_s = B * T / (t_test * 1e-3)
        kernels_ms = {"train_forward": t_fwd, "adjoint": t_bwd, "exact_test_forward": t_test,
                      "reduce_partials": t_red}
        traffic = prof.get("adjoint_traffic_bytes") if prof.get("batch") == B else None
        roofline = {"bound": "hbm",
                    "kernel": ("rk_spec_kernel" if planner.train_bwd.specialized
                               else "rk_interp_kernel") + " (adjoint program)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes": bytes_bwd,
                    "bytes_per_rollout_step": 4 * (fw.tape_words + N),
                    "issue_slot_utilisation": prof.get("adjoint_issue_active"),
                    "sass_per_warp_step": prof.get("adjoint_sass_per_warp_step"),
                    "note": prof.get("note", "")}
        roofline_other = {"train_forward_GBs": bytes_fwd / (t_fwd * 1e-3) / 1e9,
                          "exact_test_forward_GBs": bytes_test / (t_test * 1e-3) / 1e9}
        n_launch = 9 + (1 if planner.plan.needs_concurrency_projection else 0)
        bytes_note = f"tape+noise+logs ({(bytes_fwd + bytes_bwd) / 1e6:.0f} MB/iter)"
    exchange = "none" if world == 1 else ("peer" if planner._peer_ar is not None else "nccl")
    cfg = workload_config(name, n_gpus)      # identical to the reference arm's config
    if ctx.batch:
        cfg.update({"global_batch": n_gpus * B, "batch_per_gpu": B})
    setup = {"state_words": S, "action_words": A, "noise_words": N, "exchange": exchange,
             "timing": f"{bytes_note} streamed through HBM/L2 every iteration; per-kernel "
                       "numbers taken with a 256 MB L2 flush between launches",
             "cuda_graph": bool(planner._graph is not None)}
    out = {
        "name": name, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus,
        "steps": steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "setup": setup, "iters_per_sec": 1e3 / ms_per_step,
        "forward_rollout_steps_per_sec": n_gpus * test_steps_per_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "iters_per_sec": e2e_value / (n_gpus * B * T)},
        "gpu_launches": n_launch * planner.parallel_updates * steps,
        "kernels_ms": kernels_ms, "roofline": roofline, "roofline_other": roofline_other,
        "program": {"specialized": bool(planner.train_bwd.specialized),
                    "plan": "drp" if planner.is_drp else "slp", "trainable_words": planner.n_params,
                    "fwd_step_ins": fw.n_ins[1], "bwd_step_ins": bw.n_ins[1],
                    "test_step_ins": te.n_ins[1], "rpw": bw.rpw, "lanes": bw.lanes,
                    "warp_words": bw.warp_words},
    }
    if with_cpu:
        out["cpu_baseline"] = cpu_leg(name, 60, 1, budget_s=8.0, n_gpus=n_gpus)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="rollouts per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the other four workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stdout_fd = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        # keep stdout to the one JSON line: native libraries (NCCL's version banner when the
        # symmetric-memory allocator brings up its communicator) write to fd 1 directly, so fd 1
        # points at stderr until the result is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    ctx = Ctx()
    ctx.dev, ctx.world, ctx.rank, ctx.batch, ctx.keep = dev, world, rank, args.batch, []
    ctx.peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            ctx.peaks = json.load(f)
    except Exception:
        pass
    ctx.hbm_peak = float(ctx.peaks.get("hbm_gbs", 6650.0))
    ctx.peak_src = "measured (MEASURED_PEAKS.json)" if ctx.peaks else "fallback (B200_PROFILING.md)"
    # L2 hygiene: the per-kernel pass flushes L2 (256 MB write) between launches; in the
    # whole-iteration loop the tape + noise + logs written by one kernel and read by the next
    # are streamed through HBM/L2 as they would be in production (the working set is in config)
    ctx.flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    head_name = args.workload or HEADLINE
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    # the clock sampler runs over the headline's burst + timed loops: a timed region of a few
    # sub-millisecond iterations alone is shorter than nvidia-smi's sampling period
    sampler = ClockSampler(local_rank)
    sampler.start()
    line = measure_workload(head_name, ctx, args.steps, args.warmup, burst_s=2.5, with_cpu=want_cpu)
    line["clocks"] = sampler.stop()
    line.pop("name")
    if not args.workload and not args.headline_only:
        rows = []
        for name in ORDER:
            if name == head_name:
                r = {k: line[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "setup",
                                          "e2e", "roofline", "kernels_ms", "iters_per_sec",
                                          "gpu_launches")}
                if "cpu_baseline" in line:
                    r["cpu_baseline"] = line["cpu_baseline"]
                r["name"] = name
            else:
                r = measure_workload(name, ctx, args.steps, args.warmup, with_cpu=want_cpu)
                for k in ("metric", "higher_is_better", "vs_baseline", "dtype", "data", "n_gpus",
                          "steps", "warmup", "roofline_other", "forward_rollout_steps_per_sec"):
                    r.pop(k, None)
            rows.append(r)
        line["workloads"] = rows
        line["gpu_launches"] = int(sum(r["gpu_launches"] for r in rows))
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: destroying the process group while the captured
        # iteration graph (which holds the all-reduce) is alive can block forever
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
