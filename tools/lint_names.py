"""Undefined-global check without third-party linters: every LOAD_GLOBAL / LOAD_NAME in a file must be bound at module
level, be a builtin, or be bound by a `global` statement.  Usage: python tools/lint_names.py [files...] (default: repo)."""
import ast
import builtins
import dis
import os
import sys
import types


def module_bindings(tree):
    names = set()
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(node.name)
        elif isinstance(node, ast.Global):
            names.update(node.names)
    for node in tree.body:                                  # module-level assignments / loops / with / try targets
        for sub in ast.walk(node):
            if isinstance(sub, ast.Name) and isinstance(sub.ctx, (ast.Store, ast.Del)):
                names.add(sub.id)
    return names


def check(path):
    src = open(path).read()
    tree = ast.parse(src, path)
    bound = module_bindings(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__annotations__"}
    bad = []

    def walk(code):
        for ins in dis.get_instructions(code):
            if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME") and ins.argval not in bound:
                if ins.opname == "LOAD_NAME" and code.co_name != "<module>":
                    continue                                # class bodies: names bound locally in the class
                bad.append((ins.argval, code.co_name, ins.positions.lineno if ins.positions else code.co_firstlineno))
        for c in code.co_consts:
            if isinstance(c, types.CodeType):
                walk(c)

    walk(compile(src, path, "exec"))
    return bad


def check_repo_attrs(path):
    """`alias.attr` where alias is an import of one of the repo's own modules: the attribute must exist."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    tree = ast.parse(open(path).read(), path)
    alias = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.level == 0 and \
                node.module.split(".")[0] in ("hallo_b200", "oracle"):
            for a in node.names:
                alias[a.asname or a.name] = (node.module, a.name)
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name.split(".")[0] in ("hallo_b200", "oracle") and a.asname:
                    alias[a.asname] = (a.name, None)
    bad = []
    cache = {}

    def resolve(mod, name):
        key = (mod, name)
        if key not in cache:
            try:
                m = importlib.import_module(mod)
                if name is None:
                    cache[key] = m
                elif hasattr(m, name):
                    cache[key] = getattr(m, name)
                else:
                    cache[key] = importlib.import_module(mod + "." + name)
            except Exception as e:           # noqa: BLE001
                cache[key] = e
        return cache[key]

    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in alias:
            obj = resolve(*alias[node.value.id])
            if isinstance(obj, Exception):
                bad.append((f"{node.value.id} (import failed: {obj})", "<import>", node.lineno))
            elif isinstance(obj, types.ModuleType) and not hasattr(obj, node.attr):
                bad.append((f"{node.value.id}.{node.attr}", "<attribute>", node.lineno))
    return bad


def main():
    files = sys.argv[1:]
    if not files:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for d, _, fs in os.walk(root):
            if any(p in d for p in ("/.git", "/gpurun_out", "/oracle/_ref", "/baseline", "__pycache__")):
                continue
            files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
    n = 0
    for f in sorted(files):
        for name, fn, line in check(f) + check_repo_attrs(f):
            print(f"{f}:{line}: undefined name '{name}' in {fn}")
            n += 1
    print(f"{n} problem(s)")
    return 1 if n else 0


if __name__ == "__main__":
    sys.exit(main())
