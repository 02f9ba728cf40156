"""Generates tests/golden/*.npz: frozen inputs and ORACLE outputs for the three fixture domains.

    python -m tests.golden.make_golden

The reference cannot be run here (jax / pyRDDLGym are not installable offline), so
these vectors come from the CPU oracle restatement (oracle/), not from the reference: they pin
the oracle against silent drift and give the CUDA path a fixed target.  Inputs (plan parameters,
supplied noise) are stored in full so the files do not depend on torch's RNG streams.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.helpers import pack_params          # noqa: E402
from tests.parity import forward_case, gradient_case   # noqa: E402
from oracle import planner as OP               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # name, key, B, T, relaxed, gamma, compiler kwargs
    ("quadcopter_exact", "quadcopter", 5, 6, False, 1.0, {}),
    ("quadcopter_relaxed", "quadcopter", 5, 6, True, 1.0, {}),
    ("reservoir_exact", "reservoir", 6, 5, False, 1.0, {}),
    ("reservoir_relaxed", "reservoir", 6, 5, True, 0.9, {}),
    ("wildfire_exact", "wildfire", 7, 5, False, 1.0, {}),
    ("wildfire_relaxed", "wildfire", 7, 5, True, 1.0, {}),
    ("marsrover_exact", "marsrover", 6, 6, False, 1.0, {}),
    ("marsrover_relaxed", "marsrover", 6, 6, True, 1.0, {}),
    ("powergen_exact", "powergen", 6, 6, False, 1.0, {}),
    ("powergen_relaxed", "powergen", 6, 6, True, 0.95, {}),
    ("discrete_exact", "discrete", 6, 5, False, 1.0, {}),
    ("discrete_relaxed", "discrete", 6, 5, True, 1.0, {}),
    ("distributions_exact", "distributions", 6, 4, False, 1.0, {}),
    ("distributions_relaxed", "distributions", 6, 4, True, 1.0, {}),
    ("ops_exact", "ops", 6, 5, False, 1.0, {}),
    ("ops_relaxed", "ops", 6, 5, True, 0.9, {}),
    ("samplers_exact", "samplers", 6, 5, False, 1.0, {}),
    ("samplers_relaxed", "samplers", 6, 5, True, 1.0, {}),
    ("matrix_exact", "matrix", 6, 5, False, 1.0, {}),
    ("matrix_relaxed", "matrix", 6, 5, True, 0.9, {}),
    ("wildfire_relaxed_w100", "wildfire", 4, 4, True, 1.0,
     {"sigmoid_weight": 100.0, "bernoulli_sigmoid_weight": 100.0}),
]


def pack_final(c):
    out = {}
    for s in c["model"].state_fluents:
        a = c["ref"]["final"][s].detach()
        out["final_" + s] = a.reshape(a.shape[0], -1).to(
            a.dtype if a.dtype.is_floating_point else __import__("torch").int32).numpy()
    return out


def main(only=()):
    for name, key, B, T, relaxed, gamma, ckw in CASES:
        if only and name not in only:
            continue
        if relaxed:
            c = gradient_case("emu", key, B, T, gamma=gamma, **ckw)
        else:
            c = forward_case("emu", key, B, T, False, gamma=gamma)
        ref = c["ref"]
        data = {"B": B, "T": T, "gamma": gamma, "relaxed": relaxed,
                "noise": c["noise"].numpy(),
                "reward": ref["reward"].detach().numpy(),                # [B, T]
                "returns": OP.returns_from_rewards(ref["reward"].detach(), gamma).numpy(),
                "error": ref["error"].numpy().astype(np.uint32)}
        for k, v in c["params"].items():
            data["param_" + k] = v.detach().numpy()
        data.update(pack_final(c))
        if relaxed:
            data["grad"] = c["grad_ref"]
            data["loss"] = np.float32(c["loss_ref"])
        for k, v in ckw.items():
            data["ckw_" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: np.asarray(v).shape for k, v in data.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main(tuple(sys.argv[1:]))      # optional: names of the cases to (re)generate
