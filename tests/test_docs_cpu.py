"""Keeps the documents honest: every test file / test function COVERAGE.md and DESIGN.md cite exists,
and every C entry point they name is declared in include/rk.h."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(name):
    with open(os.path.join(ROOT, name)) as f:
        return f.read()


def _test_sources():
    out = {}
    for f in os.listdir(os.path.join(ROOT, "tests")):
        if f.startswith("test_") and f.endswith(".py"):
            out[f] = _read(os.path.join("tests", f))
    return out


def test_cited_tests_exist():
    sources = _test_sources()
    everything = "\n".join(sources.values())
    for doc in ("COVERAGE.md", "DESIGN.md"):
        text = _read(doc)
        for fname in set(re.findall(r"\b(test_[a-z0-9_]+\.py)\b", text)):
            assert fname in sources, f"{doc} cites tests/{fname}, which does not exist"
        for fname, func in set(re.findall(r"\b(test_[a-z0-9_]+\.py)::(test_[A-Za-z0-9_]+)", text)):
            assert re.search(rf"def {func}\b", sources[fname]), f"{doc}: {fname}::{func} not found"
        reference_symbols = {"test_loss", "test_policy", "test_compiled", "test_rollouts", "test_key",
                             "test_return", "test_log", "test_rolling_window"}
        for func in set(re.findall(r"`(test_[a-z0-9_]+)`", text)) - reference_symbols:
            assert re.search(rf"def {func}\b", everything), f"{doc} cites {func}, which does not exist"


def test_cited_entry_points_are_declared():
    header = re.sub(r"/\*.*?\*/", "", _read(os.path.join("include", "rk.h")), flags=re.S)
    declared = set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", header))
    kernels_and_internal = re.compile(r"_kernel$|^rk_(act|device|isa|launch|interp|spec|common)")
    for doc in ("COVERAGE.md", "DESIGN.md", "INTEGRATION.md", "README.md"):
        for name in set(re.findall(r"`(rk_[a-z0-9_]+)`", _read(doc))):
            if kernels_and_internal.search(name):
                continue
            # "rk_foo_forward / backward" style mentions name a declared prefix
            assert name in declared or any(d.startswith(name) for d in declared), \
                f"{doc} names {name}, which include/rk.h does not declare"


def test_integration_stub_struct_matches_header():
    """The ctypes stub a reference maintainer would paste declares ``RkSizes`` with the
    same fields, in the same order, as ``struct rk_sizes`` of include/rk.h and as the package's own
    binding -- a shorter struct makes rk_program_query write past its end."""
    header = _read(os.path.join("include", "rk.h"))
    body = re.search(r"typedef struct rk_sizes \{(.*?)\} rk_sizes;", header, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    want = []
    for decl in re.findall(r"int32_t\s+([^;]+);", body):
        want += [n.strip() for n in decl.split(",")]
    stub = re.search(r"class RkSizes\(ctypes\.Structure\):.*?_fields_ = \[.*?for n in \((.*?)\)\]",
                     _read("INTEGRATION.md"), flags=re.S).group(1)
    got = re.findall(r'"([A-Za-z_0-9]+)"', stub)
    assert got == want, (got, want)
    from pyrddlgym_jax_b200.core.engine import RkSizes
    assert [n for n, _ in RkSizes._fields_] == want
