"""Pre-compiles the specialised kernels of the bench workloads (csrc/jit/*.cubin) so that a
fresh GPU box does not pay nvcc time (the cubins are git-ignored but travel with the repo
snapshot)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyrddlgym_jax_b200.core.codegen import compile_cubin  # noqa: E402
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler  # noqa: E402
from pyrddlgym_jax_b200.core.planner import JaxDeepReactivePolicy, JaxStraightLinePlan  # noqa: E402


def prebuild_workload(name: str, verbose: bool = True):
    w = bench.WORKLOADS[name]
    model = bench.workload_model(name)
    comp = bench.make_compiler(name, model)
    exact = JaxRDDLCompiler(model)
    drp = w.get("plan") == "drp"
    plan = (JaxDeepReactivePolicy if drp else JaxStraightLinePlan)(**w["plan_kwargs"])
    plan.compile(comp, exact, {}, w["T"])
    gamma = float(model.discount)
    # a strong-scaling workload runs at global batch / N rollouts per GPU: the batch hint may change
    # the lanes-per-rollout choice, so every N of the scaling run gets its kernels
    batches = sorted({bench.per_gpu_batch(name, n) for n in (1, 2, 4, 8)})
    for B in batches:
        tb = comp.builder(plan.spec, True, (), B)
        if drp:     # the programs JaxBackpropPlanner._compile builds for a closed-loop plan
            train = (tb.forward(write_tape=False, gamma=gamma),
                     tb.backward(gamma=gamma, grad_s0=True, state_ct_in=True))
        else:
            train = (tb.backward(gamma=gamma), tb.forward(write_tape=True, gamma=gamma))
        for prog in (*train,
                     exact.builder(plan.spec, False, (), B).forward(write_tape=False, gamma=gamma)):
            path = compile_cubin(prog)[0]
            if verbose:
                print(name, B, os.path.basename(path))


def prebuild_all(verbose: bool = True):
    for name in bench.WORKLOADS:
        prebuild_workload(name, verbose)


if __name__ == "__main__":
    prebuild_all()
