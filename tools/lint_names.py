"""Undefined-name check for the host code (no third-party linter is installed in the image).

The GPU-side host paths (bench.py's GPU arms, engine.py, rcnn.py) cannot be executed in the authoring container; this
catches the cheapest class of error they could hide: a name that is never bound in the function, its enclosing
functions, the module or builtins.  Used by tests/test_host_cpu.py::test_no_undefined_names.

  python tools/lint_names.py [files...]
"""
import ast
import builtins
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bound_in(node):
    """Names bound directly in `node`'s own scope (not in nested function / class scopes)."""
    out = set()
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = node.args
        for x in a.posonlyargs + a.args + a.kwonlyargs:
            out.add(x.arg)
        if a.vararg:
            out.add(a.vararg.arg)
        if a.kwarg:
            out.add(a.kwarg.arg)
    stack = list(ast.iter_child_nodes(node)) if not isinstance(node, ast.Lambda) else [node.body]
    while stack:
        x = stack.pop()
        if isinstance(x, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(x.name)
            stack.extend(x.decorator_list)
            continue                      # new scope: its bindings are not ours
        if isinstance(x, ast.Lambda):
            continue
        if isinstance(x, ast.Name) and isinstance(x.ctx, (ast.Store, ast.Del)):
            out.add(x.id)
        elif isinstance(x, (ast.Import, ast.ImportFrom)):
            for al in x.names:
                out.add((al.asname or al.name).split(".")[0])
        elif isinstance(x, ast.ExceptHandler) and x.name:
            out.add(x.name)
        elif isinstance(x, (ast.Global, ast.Nonlocal)):
            out.update(x.names)
        elif isinstance(x, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
            for g in x.generators:        # comprehension targets: treated as bound in the enclosing scope (conservative)
                for t in ast.walk(g.target):
                    if isinstance(t, ast.Name):
                        out.add(t.id)
        stack.extend(ast.iter_child_nodes(x))
    return out


def undefined_names(path):
    tree = ast.parse(open(path).read(), path)
    problems = []

    def visit(scope, env):
        env = env | _bound_in(scope)
        body = [scope.body] if isinstance(scope, ast.Lambda) else list(ast.iter_child_nodes(scope))
        stack = list(body)
        while stack:
            x = stack.pop()
            if isinstance(x, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                if not isinstance(x, ast.Lambda):
                    stack.extend(x.decorator_list)
                    stack.extend(x.args.defaults + [d for d in x.args.kw_defaults if d is not None])
                visit(x, env)
                continue
            if isinstance(x, ast.ClassDef):
                stack.extend(x.decorator_list + x.bases)
                visit(x, env)   # class-level names are visible to the class body only; methods see them via self (not checked)
                continue
            if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Load) and x.id not in env:
                problems.append((x.lineno, x.id))
            stack.extend(ast.iter_child_nodes(x))

    visit(tree, set(dir(builtins)) | {"__file__", "__name__", "__doc__"})
    return sorted(set(problems))


def default_files():
    files = [os.path.join(REPO, "bench.py"), os.path.join(REPO, "__graft_entry__.py")]
    for root in ("siammot_b200", "oracle", "tools", "tests"):
        for d, _, names in os.walk(os.path.join(REPO, root)):
            files += [os.path.join(d, n) for n in names if n.endswith(".py")]
    return sorted(files)


if __name__ == "__main__":
    bad = 0
    for f in (sys.argv[1:] or default_files()):
        for line, name in undefined_names(f):
            print("%s:%d: undefined name %r" % (os.path.relpath(f, REPO), line, name))
            bad += 1
    sys.exit(1 if bad else 0)
