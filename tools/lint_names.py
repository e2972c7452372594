"""Tiny undefined-name check (no pyflakes in the image): flags names that are loaded in a function but bound nowhere
in its enclosing scopes / the module / builtins.  Usage: python tools/lint_names.py file.py ..."""
import ast
import builtins
import sys


def bound_names(node):
    names = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            names.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(n.name)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.arg):
            names.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            names.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            names.update(n.names)
    return names


def check(path):
    tree = ast.parse(open(path).read(), path)
    known = set(dir(builtins)) | bound_names(tree) | {"__file__", "__name__"}
    bad = []
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in known:
            bad.append((n.lineno, n.id))
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for line, name in check(p):
            print(f"{p}:{line}: undefined name {name}")
            rc = 1
    sys.exit(rc)
