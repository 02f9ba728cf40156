"""Static hygiene of the GPU-only Python paths, which the CPU suite cannot execute: every name a
function loads resolves to a local, a closure variable, a module-level name or a builtin (catches
the NameErrors that would otherwise first show up on the B200 box)."""
import ast
import builtins
import glob
import os
def check(path):
    src=open(path).read(); tree=ast.parse(src)
    mod_names=set(dir(builtins))
    for node in ast.walk(tree):
        if isinstance(node,(ast.Import,ast.ImportFrom)):
            for a in node.names: mod_names.add((a.asname or a.name).split('.')[0])
        elif isinstance(node,(ast.FunctionDef,ast.ClassDef,ast.AsyncFunctionDef)):
            mod_names.add(node.name)
    for node in tree.body:
        for t in ast.walk(node) if isinstance(node,(ast.Assign,ast.AnnAssign,ast.AugAssign,ast.For,ast.With,ast.Try,ast.If)) else []:
            if isinstance(t,ast.Name) and isinstance(t.ctx,ast.Store): mod_names.add(t.id)
    problems=[]
    def visit(fn, outer):
        local=set(outer)
        args=fn.args
        for a in args.args+args.kwonlyargs+args.posonlyargs: local.add(a.arg)
        if args.vararg: local.add(args.vararg.arg)
        if args.kwarg: local.add(args.kwarg.arg)
        for n in ast.walk(fn):
            if isinstance(n,ast.Name) and isinstance(n.ctx,(ast.Store,ast.Del)): local.add(n.id)
            elif isinstance(n,(ast.FunctionDef,ast.ClassDef,ast.AsyncFunctionDef)) and n is not fn: local.add(n.name)
            elif isinstance(n,(ast.Import,ast.ImportFrom)):
                for a in n.names: local.add((a.asname or a.name).split('.')[0])
            elif isinstance(n,ast.ExceptHandler) and n.name: local.add(n.name)
            elif isinstance(n,ast.arg): local.add(n.arg)
            elif isinstance(n,ast.comprehension):
                for t in ast.walk(n.target):
                    if isinstance(t,ast.Name): local.add(t.id)
        for n in ast.walk(fn):
            if isinstance(n,ast.Name) and isinstance(n.ctx,ast.Load) and n.id not in local and n.id not in mod_names:
                problems.append((path,n.lineno,n.id))
    for node in tree.body:
        if isinstance(node,(ast.FunctionDef,ast.AsyncFunctionDef)):
            visit(node,set())
        elif isinstance(node,ast.ClassDef):
            cls_names={n.name for n in node.body if isinstance(n,(ast.FunctionDef,ast.ClassDef))}
            for t in node.body:
                for x in ast.walk(t) if isinstance(t,(ast.Assign,ast.AnnAssign)) else []:
                    if isinstance(x,ast.Name) and isinstance(x.ctx,ast.Store): cls_names.add(x.id)
            for n in node.body:
                if isinstance(n,(ast.FunctionDef,ast.AsyncFunctionDef)):
                    visit(n,set())
    return problems


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_names_in_the_package():
    files = []
    for pat in ("pyrddlgym_jax_b200/core/*.py", "pyrddlgym_jax_b200/rddl/*.py", "pyrddlgym_jax_b200/*.py",
                "oracle/*.py", "tools/*.py", "bench.py", "__graft_entry__.py"):
        files += glob.glob(os.path.join(ROOT, pat))
    assert len(files) > 20
    problems = [pr for f in sorted(files) for pr in sorted(set(check(f)))]
    assert not problems, problems
