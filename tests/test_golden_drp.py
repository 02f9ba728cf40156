"""Committed golden vectors of the deep-reactive-policy restatement (tests/golden/drp/*.npz, made by
tests/golden/make_golden_drp.py): the oracle still reproduces them (drift guard), and the product's
host-side test policy (controllers' get_action path) matches the frozen test actions."""
import glob
import os

import numpy as np
import pytest
import torch

from .golden.make_golden_drp import CASES, build, evaluate

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "drp", "*.npz")))


def test_golden_drp_files_exist():
    assert sorted(os.path.basename(f)[:-4] for f in FILES) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_drp_matches_golden(name):
    g = np.load(os.path.join(HERE, "golden", "drp", name + ".npz"))
    m, plan, okw = build(name)
    flat = torch.from_numpy(g["flat"])
    states = {v: torch.from_numpy(g["state_" + v]) for v in m.state_fluents}
    noise = {v: torch.from_numpy(g["noise_" + v]) for v in m.action_fluents}
    acts, ent, test, grad = evaluate(m, plan, okw, flat, states, noise, int(g["step"]))
    np.testing.assert_allclose(ent.detach().numpy(), g["entropy"], rtol=2e-6, atol=2e-6)
    for v in m.action_fluents:
        np.testing.assert_allclose(acts[v].detach().numpy(), g["train_" + v], rtol=2e-6, atol=2e-6)
        if g["test_" + v].dtype.kind in "biu":
            assert np.array_equal(test[v].numpy(), g["test_" + v]), v       # bool / int: bit-exact
        else:
            np.testing.assert_allclose(test[v].numpy(), g["test_" + v], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(grad.numpy(), g["grad"], rtol=2e-5, atol=2e-5 * np.abs(g["grad"]).max())


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_test_policy_matches_golden(name):
    g = np.load(os.path.join(HERE, "golden", "drp", name + ".npz"))
    m, plan, _ = build(name)
    flat = torch.from_numpy(g["flat"])
    for b in range(int(g["B"])):
        vec = np.concatenate([g["state_" + v][b].reshape(-1).astype(np.float32)
                              for v in m.state_fluents])
        got = plan.test_actions_host(flat, vec, step=int(g["step"]))
        for v in m.action_fluents:
            want = g["test_" + v][b]
            if want.dtype.kind in "biu":
                assert np.array_equal(np.asarray(got[v]).reshape(want.shape), want), (v, b)
            else:
                np.testing.assert_allclose(np.asarray(got[v]).reshape(want.shape), want, rtol=1e-5,
                                           atol=1e-5)
