"""Much of the multi-GPU orchestration (parallel/engine.py, the gated GPU tests, the bench scripts) cannot execute on the
CPU-only authoring box; a misspelt or out-of-scope name there would only surface on a GPU.  This scan resolves every name
that is read inside a top-level function / method against what the function binds (arguments, assignments, imports, nested
definitions), the module's globals and the builtins."""
import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bound_at_module_level(tree):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for node in tree.body:
        for n in ast.walk(node) if isinstance(node, (ast.If, ast.Try, ast.With, ast.For, ast.Assign, ast.AnnAssign, ast.AugAssign)) else [node]:
            if isinstance(n, (ast.Import, ast.ImportFrom)):
                names.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)) and n is node:
                names.add(n.name)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                names.add(n.id)
    return names


def _unresolved(path):
    tree = ast.parse(open(path).read(), filename=path)
    module_names = _bound_at_module_level(tree)
    tops = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)):
            tops.append(node)
        elif isinstance(node, ast.ClassDef):
            tops += [m for m in node.body if isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef))]
    bad = []
    for fn in tops:
        local = set()
        for n in ast.walk(fn):
            if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                if not isinstance(n, ast.Lambda):
                    local.add(n.name)
                a = n.args
                local.update(x.arg for x in a.args + a.kwonlyargs + a.posonlyargs)
                local.update(x.arg for x in (a.vararg, a.kwarg) if x is not None)
            elif isinstance(n, ast.ClassDef):
                local.add(n.name)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                local.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                local.add(n.id)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                local.add(n.name)
        bad += [(fn.name, n.id, n.lineno) for n in ast.walk(fn)
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in local and n.id not in module_names]
    return bad


def test_no_unresolved_names_in_code_that_only_runs_on_gpus():
    files = (glob.glob(os.path.join(ROOT, "colearn_federated_learning_b200", "**", "*.py"), recursive=True)
             + glob.glob(os.path.join(ROOT, "scripts", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py"))
             + glob.glob(os.path.join(ROOT, "baseline", "*.py"))
             + [os.path.join(ROOT, n) for n in ("bench.py", "__graft_entry__.py", "federated_coordinator.py", "remote_worker.py")])
    problems = {os.path.relpath(p, ROOT): _unresolved(p) for p in files if os.path.exists(p)}
    problems = {k: v for k, v in problems.items() if v}
    assert not problems, problems
    assert len(files) > 80
