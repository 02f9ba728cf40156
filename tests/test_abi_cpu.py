"""CPU tests of the drop-in boundary: csrc/librk.so loads without a GPU and exports every entry
point include/rk.h declares; argument validation that needs no device; the product path
refuses to run (loudly) without a CUDA device instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rk.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pyrddlgym_jax_b200.build import build_native
    from pyrddlgym_jax_b200.core.engine import load_library
    build_native()
    return load_library()


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("rk_program_create", "rk_program_destroy", "rk_program_query",
                 "rk_rollout_forward", "rk_rollout_backward", "rk_reduce_partials",
                 "rk_optimizer_step", "rk_mean_loss", "rk_program_attach_cubin"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(lib._name)
    missing = [n for n in declared_symbols() if not hasattr(raw, n)]
    assert not missing, f"include/rk.h declares symbols librk.so does not export: {missing}"


def test_python_binding_covers_the_header(lib):
    from pyrddlgym_jax_b200.core.engine import EXPORTS
    assert sorted(EXPORTS) == declared_symbols()


def test_bad_arguments_are_rejected_without_a_device(lib):
    out = ctypes.c_void_p()
    assert lib.rk_program_create(None, 0, ctypes.byref(out)) == -1        # RK_E_BADARG
    junk = ctypes.create_string_buffer(b"\0" * 256, 256)
    assert lib.rk_program_create(junk, 256, ctypes.byref(out)) == -2      # RK_E_BADPROG (magic)
    assert lib.rk_program_query(None, None) == -1
    assert lib.rk_rollout_forward(None, None, 1, 1, *([None] * 12)) == -1
    assert lib.rk_rollout_backward(None, None, 1, 1, *([None] * 11)) == -1
    lib.rk_program_destroy(None)                                          # no-op


def test_topk_limits_match_the_binding(lib):
    from pyrddlgym_jax_b200.core import engine
    assert engine.drp_topk_limits() == (engine.TOPK_MAX_POOLED, engine.TOPK_MAX_ALLOWED)
    assert lib.rk_drp_topk_forward(None, 1, 1, 1, 1, *([None] * 5), 1, 99, 1.0, 0, None, None) == -1
    assert lib.rk_state_layernorm_forward(None, 0, 1, 1, None, None, None, 1e-6, *([None] * 4)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """The product path must fail loudly when there is no CUDA device (never run on the CPU)."""
    from pyrddlgym_jax_b200.core import engine
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
    from .helpers import fixture_model
    with pytest.raises(engine.RkError):
        JaxBackpropPlanner(fixture_model("reservoir"), JaxStraightLinePlan(), pgpe=None)
    with pytest.raises(engine.RkError):
        engine._ptr(torch.zeros(4))
