"""Pre-compiles the specialised kernels of the bench / smoke workloads (csrc/jit/*.cubin) so
that a fresh GPU box does not pay nvcc time (the cubins are git-ignored but travel with the
repo snapshot)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyrddlgym_jax_b200.core import logic  # noqa: E402
from pyrddlgym_jax_b200.core.codegen import compile_cubin  # noqa: E402
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler  # noqa: E402
from pyrddlgym_jax_b200.core.planner import JaxStraightLinePlan  # noqa: E402


def prebuild(model, B, T, **ckw):
    comp = logic.DefaultJaxRDDLCompilerWithGrad(model, **ckw)
    exact = JaxRDDLCompiler(model)
    plan = JaxStraightLinePlan()
    plan.compile(comp, exact, {}, T)
    tb = comp.builder(plan.spec, True, (), B)
    gamma = float(model.discount)
    for prog in (tb.forward(write_tape=True, gamma=gamma), tb.backward(gamma=gamma),
                 exact.builder(plan.spec, False, (), B).forward(write_tape=False, gamma=gamma)):
        print(compile_cubin(prog)[0])


if __name__ == "__main__":
    prebuild(bench.quadcopter_model(), bench.B_PER_GPU, bench.HORIZON, sigmoid_weight=10.)
