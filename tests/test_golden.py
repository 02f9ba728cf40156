"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py):
CPU: the oracle still reproduces them and the lowered programs (ISA emulator) match them;
GPU: both CUDA back-ends, through the C ABI, match them.  Tolerances as in tests/parity.py."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import planner as OP

from .parity import FLOOR_FWD, assert_close, forward_case, gradient_case

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def run_case(backend, path):
    g = np.load(path)
    key = os.path.basename(path).split("_")[0]
    B, T, gamma, relaxed = int(g["B"]), int(g["T"]), float(g["gamma"]), bool(g["relaxed"])
    ckw = {k[4:]: float(g[k]) for k in g.files if k.startswith("ckw_")}
    if relaxed:
        c = gradient_case(backend, key, B, T, gamma=gamma, given=g, **ckw)
    else:
        c = forward_case(backend, key, B, T, False, gamma=gamma, given=g)
    return g, c


def check_against_golden(g, c, rtol, gtol, oracle_side):
    """oracle_side=True compares the oracle's own outputs, False the program's."""
    relaxed = bool(g["relaxed"])
    gold_r = g["reward"]                                        # [B, T]
    scale = max(1.0, float(np.abs(gold_r).max()))
    if oracle_side:
        r = c["ref"]["reward"].detach().numpy()
        ret = OP.returns_from_rewards(c["ref"]["reward"].detach(), float(g["gamma"])).numpy()
        err = c["ref"]["error"].numpy().astype(np.uint32)
        final = {s: c["ref"]["final"][s].detach().reshape(int(g["B"]), -1)
                 for s in c["model"].state_fluents}
        grad = c.get("grad_ref")
    else:
        r = c["got"]["reward"].T
        ret = c["got"]["returns"]
        err = c["got"]["err"].T
        final = {s: c["final"][s].reshape(int(g["B"]), -1) for s in c["model"].state_fluents}
        grad = c.get("grad")
    assert_close(r, gold_r, rtol, "reward", FLOOR_FWD, scale_min=1.0)
    assert_close(ret, g["returns"], rtol, "returns", FLOOR_FWD, scale_min=scale)
    np.testing.assert_array_equal(err, g["error"])
    for s, v in final.items():
        gold = g["final_" + s]
        if gold.dtype.kind in "iu" or v.dtype in (torch.bool, torch.int32):
            assert np.array_equal(v.to(torch.int32).numpy(), gold.astype(np.int32)), \
                f"{s}: bool/int fluent not bit-exact"
        else:
            assert_close(v.numpy(), gold, rtol, f"final state {s}", FLOOR_FWD, scale_min=1.0)
    if relaxed:
        assert_close(grad, g["grad"], gtol, "gradient")


def test_golden_files_exist():
    assert len(FILES) >= 7


@pytest.mark.parametrize("name", NAMES)
def test_oracle_and_emulator_match_golden(name):
    g, c = run_case("emu", os.path.join(HERE, "golden", name + ".npz"))
    check_against_golden(g, c, rtol=2e-6, gtol=2e-5, oracle_side=True)    # oracle drift
    check_against_golden(g, c, rtol=1e-5, gtol=1e-4, oracle_side=False)   # lowered program


@pytest.mark.gpu
@pytest.mark.parametrize("backend", ["cuda", "jit"])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(cuda_device, backend, name):
    g, c = run_case(backend, os.path.join(HERE, "golden", name + ".npz"))
    check_against_golden(g, c, rtol=1e-5, gtol=1e-4, oracle_side=False)
