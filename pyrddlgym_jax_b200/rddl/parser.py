"""A compact RDDL parser (domain / non-fluents / instance blocks).

The reference never parses RDDL itself: it receives an already-lifted model from pyRDDLGym
(compiler.py:82, planner.py:2399).  pyRDDLGym is not available to this build, so this module
owns the subset of the language the planner's hot path consumes: types (objects and enums),
pvariables, cpfs, reward, constraints, and the expression families dispatched in
compiler.py:936-964.  Operator precedence follows the published RDDL grammar
(if / aggregations lowest, then <=>, =>, |, ^, ~, comparisons, +-, */, unary minus).
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Tuple

from .expr import Expr


class RDDLParseError(Exception):
    pass


_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<num>(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?)
  | (?P<var>\?[A-Za-z_][A-Za-z0-9_\-]*)
  | (?P<enum>@[A-Za-z0-9_][A-Za-z0-9_\-]*)
  | (?P<obj>\$[A-Za-z0-9_][A-Za-z0-9_\-]*)
  | (?P<id>[A-Za-z][A-Za-z0-9_\-]*'?)
  | (?P<op><=>|=>|<=|>=|==|~=|[-+*/<>~^&|=(){}\[\]:;,.])
""", re.X)

AGGREGATIONS = {
    "sum": "sum", "avg": "avg", "prod": "prod", "min": "minimum", "max": "maximum",
    "forall": "forall", "exists": "exists", "argmin": "argmin", "argmax": "argmax",
}
FUNCTIONS = {
    "abs", "sgn", "round", "floor", "ceil", "cos", "sin", "tan", "acos", "asin", "atan",
    "cosh", "sinh", "tanh", "exp", "ln", "sqrt", "lngamma", "gamma",
    "div", "mod", "fmod", "min", "max", "pow", "log", "hypot",
}
MATRIX_OPS = {"inverse", "pinverse", "cholesky"}
RANDOM_VECTORS = {"MultivariateNormal", "MultivariateStudent", "Dirichlet", "Multinomial"}
DISTRIBUTIONS = {
    "KronDelta", "DiracDelta", "Uniform", "Bernoulli", "Normal", "Poisson", "Exponential",
    "Weibull", "Gamma", "Binomial", "NegativeBinomial", "Beta", "Geometric", "Pareto",
    "Student", "Gumbel", "Laplace", "Cauchy", "Gompertz", "ChiSquare", "Kumaraswamy",
    "Discrete", "UnnormDiscrete",
}

# binary operator precedence (higher binds tighter)
_BINARY = {
    "<=>": (1, "boolean"), "=>": (2, "boolean"), "|": (3, "boolean"),
    "^": (4, "boolean"), "&": (4, "boolean"),
    "~": (4, "boolean"),          # binary form = xor (compiler.py:1233-1274: '~' with two arguments)
    "==": (6, "relational"), "~=": (6, "relational"), "<": (6, "relational"),
    "<=": (6, "relational"), ">": (6, "relational"), ">=": (6, "relational"),
    "+": (7, "arithmetic"), "-": (7, "arithmetic"),
    "*": (8, "arithmetic"), "/": (8, "arithmetic"),
}
_PREC_NOT = 5
_PREC_UMINUS = 9


def tokenize(text: str) -> List[Tuple[str, str, int]]:
    text = re.sub(r"/\*.*?\*/", lambda m: " " * len(m.group(0)), text, flags=re.S)
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN_RE.match(text, pos)
        if m is None:
            line = text.count("\n", 0, pos) + 1
            raise RDDLParseError(f"unexpected character {text[pos]!r} at line {line}")
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group(kind), pos))
        pos = m.end()
    out.append(("eof", "", pos))
    return out


class _Parser:
    def __init__(self, text: str):
        self.text = text
        self.toks = tokenize(text)
        self.i = 0

    # -- token helpers ---------------------------------------------------------------
    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def error(self, msg):
        _, val, pos = self.peek()
        line = self.text.count("\n", 0, pos) + 1
        raise RDDLParseError(f"{msg} (near {val!r}, line {line})")

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            return self.next()
        return None

    def expect(self, val):
        t = self.accept(val)
        if t is None:
            self.error(f"expected {val!r}")
        return t

    def ident(self):
        kind, val, _ = self.peek()
        if kind != "id":
            self.error("expected identifier")
        self.next()
        return val

    # -- top level ---------------------------------------------------------------------
    def parse_file(self) -> Dict[str, Any]:
        blocks = {"domain": None, "non-fluents": [], "instance": []}
        while self.peek()[0] != "eof":
            kw = self.ident()
            if kw == "domain":
                blocks["domain"] = self.parse_domain()
            elif kw == "non-fluents":
                blocks["non-fluents"].append(self.parse_nonfluents())
            elif kw == "instance":
                blocks["instance"].append(self.parse_instance())
            else:
                self.error(f"unknown top-level block {kw!r}")
        return blocks

    def parse_domain(self):
        dom = {"name": self.ident(), "types": {}, "pvariables": {}, "cpfs": [],
               "reward": None, "invariants": [], "preconditions": [], "terminations": [],
               "requirements": []}
        self.expect("{")
        while not self.accept("}"):
            kw = self.ident()
            if kw == "requirements":
                self.expect("=")
                self.expect("{")
                while not self.accept("}"):
                    dom["requirements"].append(self.ident())
                    self.accept(",")
                self.expect(";")
            elif kw == "types":
                self.expect("{")
                while not self.accept("}"):
                    name = self.ident()
                    self.expect(":")
                    if self.accept("{"):
                        vals = []
                        while not self.accept("}"):
                            vals.append(self.next()[1])
                            self.accept(",")
                        dom["types"][name] = vals
                    else:
                        self.expect("object")
                        dom["types"][name] = None
                    self.expect(";")
                self.expect(";")
            elif kw == "pvariables":
                self.expect("{")
                while not self.accept("}"):
                    self.parse_pvariable(dom)
                self.expect(";")
            elif kw in ("cpfs", "cdfs"):
                self.expect("{")
                while not self.accept("}"):
                    name = self.ident()
                    params = []
                    if self.accept("("):
                        while not self.accept(")"):
                            params.append(self.next()[1])
                            self.accept(",")
                    self.expect("=")
                    expr = self.parse_expr()
                    self.expect(";")
                    dom["cpfs"].append((name, params, expr))
                self.expect(";")
            elif kw == "reward":
                self.expect("=")
                dom["reward"] = self.parse_expr()
                self.expect(";")
            elif kw in ("state-invariants", "action-preconditions", "termination",
                        "state-action-constraints"):
                key = {"state-invariants": "invariants", "termination": "terminations"
                       }.get(kw, "preconditions")
                self.expect("{")
                while not self.accept("}"):
                    dom[key].append(self.parse_expr())
                    self.expect(";")
                self.expect(";")
            else:
                self.error(f"unknown domain section {kw!r}")
        return dom

    def parse_pvariable(self, dom):
        name = self.ident()
        ptypes = []
        if self.accept("("):
            while not self.accept(")"):
                ptypes.append(self.ident())
                self.accept(",")
        self.expect(":")
        self.expect("{")
        kind = self.ident()
        self.expect(",")
        prange = self.ident()
        default = None
        while self.accept(","):
            key = self.ident()
            self.expect("=")
            if key == "default":
                default = self.parse_literal()
            else:  # level = k (old interm-fluent syntax)
                self.next()
        self.expect("}")
        self.expect(";")
        dom["pvariables"][name] = {"kind": kind, "params": ptypes, "range": prange,
                                   "default": default}

    def parse_literal(self):
        kind, val, _ = self.next()
        if kind == "num":
            return float(val) if re.search(r"[.eE]", val) else int(val)
        if kind == "op" and val == "-":
            v = self.parse_literal()
            return -v
        if kind == "id" and val in ("true", "false"):
            return val == "true"
        if kind == "id" and val in ("pos-inf", "neg-inf"):
            return float("inf") if val == "pos-inf" else float("-inf")
        if kind in ("enum", "obj", "id"):
            return val.lstrip("$")
        self.error("expected literal")

    def parse_objects(self):
        objs = {}
        self.expect("{")
        while not self.accept("}"):
            t = self.ident()
            self.expect(":")
            self.expect("{")
            vals = []
            while not self.accept("}"):
                vals.append(self.next()[1].lstrip("$"))
                self.accept(",")
            objs[t] = vals
            self.expect(";")
        self.expect(";")
        return objs

    def parse_assignments(self):
        """``name(o1, o2) = value;`` / ``name(o1);`` / ``~name;`` lists."""
        out = []
        self.expect("{")
        while not self.accept("}"):
            neg = self.accept("~") is not None
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    args.append(self.next()[1].lstrip("$"))
                    self.accept(",")
            if self.accept("="):
                val = self.parse_literal()
            else:
                val = not neg
            self.expect(";")
            out.append((name, args, val))
        self.expect(";")
        return out

    def parse_nonfluents(self):
        nf = {"name": self.ident(), "domain": None, "objects": {}, "values": []}
        self.expect("{")
        while not self.accept("}"):
            kw = self.ident()
            if kw == "domain":
                self.expect("=")
                nf["domain"] = self.ident()
                self.expect(";")
            elif kw == "objects":
                nf["objects"] = self.parse_objects()
            elif kw == "non-fluents":
                nf["values"] = self.parse_assignments()
            else:
                self.error(f"unknown non-fluents section {kw!r}")
        return nf

    def parse_instance(self):
        inst = {"name": self.ident(), "domain": None, "non-fluents": None, "objects": {},
                "init-state": [], "max-nondef-actions": float("inf"), "horizon": 40,
                "discount": 1.0}
        self.expect("{")
        while not self.accept("}"):
            kw = self.ident()
            if kw in ("domain", "non-fluents") and self.peek()[1] == "=":
                self.expect("=")
                inst[kw] = self.ident()
                self.expect(";")
            elif kw == "objects":
                inst["objects"] = self.parse_objects()
            elif kw == "init-state":
                inst["init-state"] = self.parse_assignments()
            elif kw in ("max-nondef-actions", "horizon", "discount"):
                self.expect("=")
                inst[kw] = self.parse_literal()
                self.expect(";")
            else:
                self.error(f"unknown instance section {kw!r}")
        return inst

    # -- expressions -----------------------------------------------------------------
    def parse_expr(self, min_prec: int = 0) -> Expr:
        lhs = self.parse_prefix()
        while True:
            kind, val, _ = self.peek()
            if kind != "op" or val not in _BINARY:
                break
            prec, family = _BINARY[val]
            if prec < min_prec:
                break
            self.next()
            rhs = self.parse_expr(prec + 1)
            # n-ary flattening for + * ^ | (compiler.py:1136-1144 folds them left anyway)
            if val in ("+", "*", "^", "&", "|") and lhs.etype == (family, val) \
                    and not getattr(lhs, "_paren", False):
                lhs = Expr((family, val), list(lhs.args) + [rhs])
            else:
                lhs = Expr((family, val), [lhs, rhs])
        return lhs

    def parse_typed_vars(self):
        tv = []
        self.expect("{")
        while not self.accept("}"):
            kind, v, _ = self.next()
            if kind != "var":
                self.error("expected ?variable")
            self.expect(":")
            tv.append((v, self.ident()))
            self.accept(",")
        return tv

    def _var(self) -> str:
        kind, v, _ = self.next()
        if kind != "var":
            self.error("expected ?variable")
        return v

    def parse_prefix(self) -> Expr:
        kind, val, _ = self.peek()
        if kind == "num":
            self.next()
            if re.search(r"[.eE]", val):
                return Expr(("constant", "real"), float(val))
            return Expr(("constant", "int"), int(val))
        if kind == "op":
            if val == "-":
                self.next()
                arg = self.parse_expr(_PREC_UMINUS)
                if arg.etype[0] == "constant" and arg.etype[1] in ("real", "int"):
                    return Expr(arg.etype, -arg.args)
                return Expr(("arithmetic", "-"), [arg])
            if val == "+":
                self.next()
                return self.parse_expr(_PREC_UMINUS)
            if val == "~":
                self.next()
                return Expr(("boolean", "~"), [self.parse_expr(_PREC_NOT)])
            if val in ("(", "["):
                self.next()
                e = self.parse_expr()
                self.expect(")" if val == "(" else "]")
                return e  # grouping: prevents n-ary flattening across the parens
            self.error("unexpected operator")
        if kind == "var":
            self.next()
            return Expr(("pvar", val), (val, None))
        if kind in ("enum", "obj"):
            self.next()
            return Expr(("pvar", val.lstrip("$@")), (val.lstrip("$@"), None))
        if kind != "id":
            self.error("unexpected token in expression")
        # identifiers
        if val in ("true", "false"):
            self.next()
            return Expr(("constant", "bool"), val == "true")
        if val == "if":
            self.next()
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            self.expect("then")
            a = self.parse_expr()
            self.expect("else")
            b = self.parse_expr()
            return Expr(("control", "if"), [c, a, b])
        if val == "switch":
            self.next()
            self.expect("(")
            pred = self.parse_expr()
            self.expect(")")
            self.expect("{")
            cases = []
            while not self.accept("}"):
                if self.accept("default"):
                    label = None
                else:
                    self.expect("case")
                    label = self.next()[1].lstrip("$@")
                self.expect(":")
                cases.append((label, self.parse_expr()))
                self.accept(",")
            return Expr(("control", "switch"), (pred, cases))
        m = re.fullmatch(r"(sum|avg|prod|min|max|forall|exists|argmin|argmax)_", val)
        if m and self.peek(1)[1] == "{":
            self.next()
            tv = self.parse_typed_vars()
            body = self.parse_expr()
            return Expr(("aggregation", AGGREGATIONS[m.group(1)]), [*tv, body])
        # matrix algebra (compiler.py:2383-2434):  det_{?r : t, ?c : t}[ expr ]  and
        # inverse / pinverse / cholesky[row=?r, col=?c][ expr ]  over two variables already in scope
        if val == "det_" and self.peek(1)[1] == "{":
            self.next()
            tv = self.parse_typed_vars()
            if len(tv) != 2:
                self.error("det_ takes two typed variables (row, column)")
            return Expr(("matrix", "det"), [*tv, self.parse_expr()])
        if val in MATRIX_OPS and self.peek(1)[1] == "[":
            self.next()
            self.expect("[")
            self.expect("row")
            self.expect("=")
            row = self._var()
            self.expect(",")
            self.expect("col")
            self.expect("=")
            col = self._var()
            self.expect("]")
            self.expect("[")
            body = self.parse_expr()
            self.expect("]")
            return Expr(("matrix", val), [(row, col), body])
        # random vectors (compiler.py:2247-2377):  Name[?i](args) samples jointly over the objects of ?i
        # (a variable of the enclosing scope); a second bracket variable names the column index of a
        # covariance argument, e.g.  MultivariateNormal[?i, ?j](mu(?i), Sigma(?i, ?j))
        if val in RANDOM_VECTORS and self.peek(1)[1] == "[":
            self.next()
            self.expect("[")
            svars = [self._var()]
            while self.accept(","):
                svars.append(self._var())
            self.expect("]")
            self.expect("(")
            args = []
            while not self.accept(")"):
                args.append(self.parse_expr())
                self.accept(",")
            return Expr(("randomvector", val), [tuple(svars), *args])
        if val in FUNCTIONS and self.peek(1)[1] == "[":
            self.next()
            self.expect("[")
            args = [self.parse_expr()]
            while self.accept(","):
                args.append(self.parse_expr())
            self.expect("]")
            return Expr(("func", val), args)
        if val in DISTRIBUTIONS and self.peek(1)[1] == "(":
            self.next()
            self.expect("(")
            if val in ("Discrete", "UnnormDiscrete"):
                etype_name = self.ident()
                cases = []
                while self.accept(","):
                    label = self.next()[1].lstrip("$@")
                    self.expect(":")
                    cases.append((label, self.parse_expr()))
                self.expect(")")
                return Expr(("randomvar", val), (etype_name, cases))
            args = []
            while not self.accept(")"):
                args.append(self.parse_expr())
                self.accept(",")
            return Expr(("randomvar", val), args)
        # plain pvariable reference
        self.next()
        params = None
        if self.accept("("):
            params = []
            while not self.accept(")"):
                k2, v2, _ = self.peek()
                if k2 == "var":
                    self.next()
                    params.append(v2)
                elif k2 in ("enum", "obj"):
                    self.next()
                    params.append(v2.lstrip("$@"))
                elif k2 == "id" and self.peek(1)[1] in (",", ")"):
                    self.next()
                    params.append(v2)  # object literal or zero-arity pvar; resolved later
                else:
                    params.append(self.parse_expr())
                self.accept(",")
        return Expr(("pvar", val), (val, params))


def parse_rddl(text: str) -> Dict[str, Any]:
    """Parse RDDL source text (domain and/or non-fluents and/or instance blocks)."""
    return _Parser(text).parse_file()
