"""Expression AST of the RDDL front-end.

The reference delegates parsing to pyRDDLGym (compiler.py:29-42) and only consumes
``expr.etype == (family, op)``, ``expr.args`` and a unique ``expr.id``
(compiler.py:936-964, logic.py:265-266).  This node keeps that surface so the lowering
and the oracle can dispatch on the same families the reference dispatches on.
"""
from __future__ import annotations

import itertools
from typing import Any, Tuple

_ids = itertools.count(1)


class Expr:
    """One node of a CPF / reward / constraint expression tree.

    etype families: constant, pvar, arithmetic, relational, boolean, aggregation,
    func, control, randomvar (the dispatch keys of compiler.py:936-964).
    """

    __slots__ = ("etype", "args", "id")

    def __init__(self, etype: Tuple[str, str], args: Any):
        self.etype = etype
        self.args = args
        self.id = next(_ids)

    def __repr__(self) -> str:
        fam, op = self.etype
        if fam == "constant":
            return repr(self.args)
        if fam == "pvar":
            name, params = self.args
            return name if not params else f"{name}({', '.join(map(str, params))})"
        if fam == "aggregation":
            *tv, body = self.args
            vs = ", ".join(f"{v}: {t}" for v, t in tv)
            return f"({op}_{{{vs}}} {body!r})"
        if fam == "control" and op == "if":
            c, a, b = self.args
            return f"(if {c!r} then {a!r} else {b!r})"
        if fam in ("func", "randomvar"):
            return f"{op}[{', '.join(map(repr, self.args))}]"
        if len(self.args) == 1:
            return f"({op}{self.args[0]!r})"
        return "(" + f" {op} ".join(map(repr, self.args)) + ")"


def walk(expr: Expr):
    """Pre-order traversal of all Expr nodes (including nested pvar arguments)."""
    yield expr
    fam, op = expr.etype
    if fam == "constant":
        return
    if fam == "pvar":
        _, params = expr.args
        for p in params or ():
            if isinstance(p, Expr):
                yield from walk(p)
        return
    if fam == "aggregation":
        yield from walk(expr.args[-1])
        return
    if fam == "control" and op == "switch":
        pred, cases = expr.args
        yield from walk(pred)
        for _, e in cases:
            yield from walk(e)
        return
    if fam == "randomvar" and op in ("Discrete", "UnnormDiscrete"):
        _, cases = expr.args
        for _, e in cases:
            yield from walk(e)
        return
    for a in expr.args:
        if isinstance(a, Expr):
            yield from walk(a)
