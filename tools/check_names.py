#!/usr/bin/env python
"""Poor man's pyflakes (none is installed in the image): reports names that a function body loads
but that are bound nowhere -- not in the function or an enclosing one, not at module level, not a
builtin.  Catches the NameError-at-run-time class of slips in scripts that only run on the GPU box
(bench.py, tools/*.py).   python tools/check_names.py bench.py tools/*.py"""
import ast
import builtins
import sys


def direct_bindings(node):
    """Names bound directly in `node`'s own scope (not inside nested functions / classes / lambdas)."""
    out = set()
    stack = list(ast.iter_child_nodes(node))
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = node.args
        for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
            out.add(x.arg)
    while stack:
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
            continue  # its body is another scope
        if isinstance(n, ast.Lambda):
            continue
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
        stack.extend(ast.iter_child_nodes(n))
    return out


def check(path):
    tree = ast.parse(open(path).read(), path)
    base = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    bad = []

    def visit(node, env):
        scope = env | direct_bindings(node)
        stack = list(ast.iter_child_nodes(node))
        while stack:
            n = stack.pop()
            if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                # decorators / defaults are evaluated in the enclosing scope
                for d in getattr(n, "decorator_list", []) + n.args.defaults + [x for x in n.args.kw_defaults if x is not None]:
                    stack.append(d)
                visit(n, scope)
                continue
            if isinstance(n, ast.ClassDef):
                visit(n, scope)
                continue
            if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                visit(n, scope)  # comprehension targets are Store names inside it
                continue
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in scope:
                bad.append((n.lineno, n.id))
            stack.extend(ast.iter_child_nodes(n))

    visit(tree, base)
    return sorted(set(bad))


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for line, name in check(p):
            print("%s:%d: undefined name %r" % (p, line, name))
            rc = 1
    sys.exit(rc)
