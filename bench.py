#!/usr/bin/env python
"""bench.py -- planner iterations of the Quadcopter SLP workload on N B200s (one node).

Workload (BASELINE.json configs[1]): Quadcopter instance 1 (4 drones, S = 48 state words,
A = 16 action words, no noise), straight-line plan, batch 4096 rollouts per GPU, horizon 100,
rmsprop lr 0.03 (examples/configs/Quadcopter_physics_slp.cfg), pgpe=None.

One bench "step" = one planner iteration as the reference runs it (planner.py:3245-3260):
relaxed forward rollouts writing the state tape, loss, backprop-through-time, gradient
reduction (+ NCCL all-reduce when N > 1), rmsprop step + box projection, then the exact-model
test rollouts.  ``value`` = train rollout-steps per second (B*T per iteration, all GPUs)
with inputs resident in HBM; ``e2e`` = the same through the public planner API with the
initial state and plan copied from pinned host memory and the losses read back every step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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

WORKLOAD = "Quadcopter instance 1, slp, batch 4096/GPU, horizon 100, fp32, rmsprop lr 0.03, pgpe=None"
B_PER_GPU = 4096
HORIZON = 100
LR = 0.03
METRIC = "rollout_steps_per_sec"
UNIT = "rollout-steps/s"


def quadcopter_model():
    from pyrddlgym_jax_b200.rddl.model import load_model
    d = os.path.join(ROOT, "pyrddlgym_jax_b200", "domains", "Quadcopter")
    return load_model(os.path.join(d, "domain.rddl"), os.path.join(d, "instance1.rddl"))


# ---------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
def oracle_iteration_fn(model, B, T, lr):
    """One planner iteration of the CPU oracle (torch-CPU fp32 restatement of the reference)."""
    from oracle.model_eval import OracleModel
    from oracle import planner as OP
    from pyrddlgym_jax_b200.core import logic
    relax = logic.DefaultJaxRDDLCompilerWithGrad(model, sigmoid_weight=10.).relax_config()
    om_train, om_test = OracleModel(model, relax), OracleModel(model, None)
    g = torch.Generator().manual_seed(42)
    params = {a: (torch.randn(T, model.variable_size(a), generator=g) * 0.01)
              for a in model.action_fluents}
    nu = {a: torch.zeros_like(p) for a, p in params.items()}
    bounds = model.bounds

    def iteration():
        P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        out = OP.rollout(om_train, lambda t, fls: OP.slp_train_actions(model, P, t), B, T, None)
        loss = OP.mean_loss(out["reward"], model.discount)
        loss.backward()
        with torch.no_grad():
            for k in params:
                p, nu[k] = OP.rmsprop_step(params[k], P[k].grad, nu[k], lr)
                lo, hi = bounds[k]
                params[k] = torch.minimum(torch.maximum(p, torch.as_tensor(lo).reshape(1, -1)),
                                          torch.as_tensor(hi).reshape(1, -1))
            test = OP.rollout(om_test, lambda t, fls: OP.slp_test_actions(
                model, {k: v.reshape(T, *model.variable_shape(k)) for k, v in params.items()},
                t, bounds), B, T, None)
            tl = OP.mean_loss(test["reward"], model.discount)
        return float(loss), float(tl)
    return iteration


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's planner iteration, all host
    threads, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = quadcopter_model()
    B, T = B_PER_GPU, HORIZON
    it = oracle_iteration_fn(model, B, T, LR)
    for _ in range(args.warmup):
        it()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        it()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = B * T / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "iters_per_sec": 1.0 / dt,
        "config": {"workload": WORKLOAD, "note": "CPU restatement (oracle/), not jax: jax, "
                   "pyRDDLGym and rddlrepository are not installable offline"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full planner iterations (B={B}, T={T}) of the "
                                   f"torch-CPU oracle, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
    from pyrddlgym_jax_b200.core import engine

    model = quadcopter_model()
    B, T = args.batch, HORIZON
    planner = JaxBackpropPlanner(
        model, JaxStraightLinePlan(), batch_size_train=B, batch_size_test=B,
        optimizer="rmsprop", optimizer_kwargs={"learning_rate": LR}, pgpe=None,
        compiler_kwargs={"sigmoid_weight": 10.}, device=dev)
    planner.initialize(42)
    fw, bw, te = planner.train_fwd.program, planner.train_bwd.program, planner.test_fwd.program
    S, A, N = fw.S, fw.A, fw.N

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graph) ------------------------------------------
    W = max(args.warmup, 3)
    for _ in range(W):
        planner.iterate()
    barrier()

    # working set: tape 4*S*B*T = 78.6 MB + reward/err/partials; L2 is flushed between timed
    # iterations of the per-kernel pass; the whole-iteration loop re-streams the tape (written
    # by the forward kernel, read by the adjoint kernel) which exceeds what stays resident
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    # ---- device-resident timed loop: EXACTLY K iterations ---------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        planner.iterate()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_local = torch.tensor([ms], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_total = float(t_local.item())
    ms_per_step = ms_total / args.steps
    value = n_gpus * B * T / (ms_per_step * 1e-3)

    # ---- e2e: public API with host buffers, H2D + D2H inside the timed region --------------
    s0_train_h = planner.s0_train.cpu().pin_memory()
    s0_test_h = planner.s0_test.cpu().pin_memory()
    params_h = planner.params.cpu().pin_memory()
    out_h = torch.empty(3, planner.parallel_updates).pin_memory()
    h2d = (s0_train_h.numel() + s0_test_h.numel() + params_h.numel()) * 4
    d2h = out_h.numel() * 4 + params_h.numel() * 4
    pback_h = torch.empty_like(params_h).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        planner.s0_train.copy_(s0_train_h, non_blocking=True)      # init_sim_state from host
        planner.s0_test.copy_(s0_test_h, non_blocking=True)
        planner.params.copy_(params_h, non_blocking=True)          # warm-start guess
        planner.iterate()
        out_h.copy_(torch.stack([planner.train_loss, planner.test_loss_buf,
                                 planner.zero_flag.float()]), non_blocking=True)
        pback_h.copy_(planner.params, non_blocking=True)           # callback['params']
        torch.cuda.synchronize()
        params_h.copy_(pback_h)
    t_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * B * T / (float(t_e2e.item()) / args.steps * 1e-3)

    # ---- per-kernel timing (CUDA events on the launch stream, L2 flushed between launches) --
    def time_kernel(fn, reps=20):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts)), float(np.min(ts))

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

    def k_opt():
        engine.reduce_partials(planner.gradp, planner.nwarps, T * A, planner.grad[0])

    t_fwd, _ = time_kernel(k_fwd)
    t_bwd, _ = time_kernel(k_bwd)
    t_test, _ = time_kernel(k_test)
    t_red, _ = time_kernel(k_opt)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"
    # algorithmic bytes of the adjoint kernel (dominant): tape + noise read once per
    # rollout-step, dL/dreturn read, per-warp gradient partial rows written
    bytes_bwd = B * T * 4 * (S + N) + 4 * B + planner.nwarps * T * A * 4
    bytes_fwd = B * T * (4 * S + 4 * N + 4 + 4) + 4 * S * B + 4 * B
    bytes_test = B * T * (4 * te.N + 4 + 4) + 4 * te.S * B * 2 + 4 * B
    achieved = bytes_bwd / (t_bwd * 1e-3) / 1e9
    test_steps_per_s = B * T / (t_test * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": n_gpus * B, "horizon": T,
                   "state_words": S, "action_words": A, "noise_words": N,
                   "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
                   "timing": "tape+partials (80 MB/iter) streamed through HBM every iteration; "
                             "per-kernel numbers taken with a 256 MB L2 flush between launches",
                   "cuda_graph": bool(planner._graph is not None)},
        "iters_per_sec": 1e3 / ms_per_step,
        "forward_rollout_steps_per_sec": n_gpus * test_steps_per_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "iters_per_sec": e2e_value / (n_gpus * B * T)},
        "gpu_launches": 9 * planner.parallel_updates * args.steps,
        "kernels_ms": {"train_forward": t_fwd, "adjoint": t_bwd, "exact_test_forward": t_test,
                       "reduce_partials": t_red},
        "roofline": {"bound": "hbm",
                     "kernel": ("rk_spec_kernel" if planner.train_bwd.specialized
                                else "rk_interp_kernel") + " (adjoint program)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes": bytes_bwd,
                     "note": "not HBM-bound by construction: 48 state words per rollout-step "
                             "carry ~290 adjoint instructions (19 sin + 19 cos + 2 tan); with "
                             "B=4096 there are only 512 warps, so the kernel is latency-bound"},
        "roofline_other": {
            "train_forward_GBs": bytes_fwd / (t_fwd * 1e-3) / 1e9,
            "exact_test_forward_GBs": bytes_test / (t_test * 1e-3) / 1e9},
        "clocks": clocks,
        "program": {"specialized": bool(planner.train_bwd.specialized),
                    "fwd_step_ins": fw.n_ins[1], "bwd_step_ins": bw.n_ins[1],
                    "test_step_ins": te.n_ins[1], "rpw": bw.rpw, "lanes": bw.lanes,
                    "warp_words": bw.warp_words},
    }

    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        it = oracle_iteration_fn(model, B, T, LR)
        it()
        n_cpu, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0 and n_cpu < 40:
            it()
            n_cpu += 1
        dt = (time.perf_counter() - t0) / n_cpu
        line["cpu_baseline"] = {"value": B * T / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "iters_per_sec": 1.0 / dt,
                                "sample": f"{n_cpu} full planner iterations (B={B}, T={T}) of the "
                                          f"torch-CPU oracle restatement, {cores} threads"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
