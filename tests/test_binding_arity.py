"""Call sites of the C++ extension take positional arguments only (the bindings define no keywords): a call with the
wrong number of arguments would only fail on a GPU.  This test compares, statically, every call of an extension function
in the package / scripts / tests with the arity pybind reports for the built modules (CUDA extension, SIMT build, host
executor — all importable on a CPU-only box)."""
import ast
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arities(mod):
    out = {}
    for name in dir(mod):
        fn = getattr(mod, name)
        doc = getattr(fn, "__doc__", None) or ""
        first = doc.splitlines()[0] if doc else ""
        if not first.startswith(name + "("):
            continue
        inner = first[len(name) + 1: first.rfind(") ->") if ") ->" in first else first.rfind(")")]
        # top-level parameters ("arg0: int", or "name: type = default" where the binding names its arguments)
        params, depth, cur = [], 0, ""
        for ch in inner:
            if ch in "[(":
                depth += 1
            elif ch in "])":
                depth -= 1
            if ch == "," and depth == 0:
                params.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            params.append(cur)
        required = sum(1 for q in params if "=" not in q)
        out[name] = (required, len(params))
    return out


@pytest.fixture(scope="module")
def arities():
    from colearn_federated_learning_b200.ops import _ext, conv, host
    mods = {}
    cuda_mod = _ext.load()
    if cuda_mod is None:
        pytest.skip("CUDA extension not built")
    mods["cuda"] = _arities(cuda_mod)
    simt = conv.load_simt()
    if simt is not None:
        mods["simt"] = _arities(simt)
    if host.available():
        mods["host"] = _arities(host.load())
    return mods


# receiver names that denote a given module at the call sites
RECEIVERS = {"cuda": {"ext", "_ext.require()", "self.ext", "self._ext", "_ext.load()"}, "simt": {"simt", "simt_mod"}, "host": {"mod"}}


def _calls(path):
    tree = ast.parse(open(path).read(), filename=path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
            try:
                recv = ast.unparse(node.func.value)
            except Exception:  # noqa: BLE001
                continue
            if any(isinstance(a, ast.Starred) for a in node.args) or node.keywords:
                continue
            yield recv, node.func.attr, len(node.args), node.lineno


def test_extension_call_sites_pass_the_number_of_arguments_the_bindings_take(arities):
    files = (glob.glob(os.path.join(ROOT, "colearn_federated_learning_b200", "**", "*.py"), recursive=True)
             + glob.glob(os.path.join(ROOT, "scripts", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py"))
             + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    checked, bad = 0, []
    for path in files:
        for recv, fn, nargs, line in _calls(path):
            for kind, names in RECEIVERS.items():
                if kind == "host" and not path.endswith(os.path.join("ops", "host.py")):
                    continue
                if recv in names and kind in arities and fn in arities[kind]:
                    checked += 1
                    lo, hi = arities[kind][fn]
                    if not lo <= nargs <= hi:
                        bad.append(f"{os.path.relpath(path, ROOT)}:{line}: {recv}.{fn}() passes {nargs} args, the {kind} binding takes {lo}..{hi}")
    # `mod.<fn>(...)` in ops/conv.py addresses the CUDA extension or the SIMT build depending on the branch: either arity
    for path in files:
        if path.endswith(os.path.join("ops", "host.py")):
            continue
        for recv, fn, nargs, line in _calls(path):
            if recv in ("mod", "self.mod"):
                want = [arities[k][fn] for k in ("cuda", "simt") if k in arities and fn in arities[k]]
                if want:
                    checked += 1
                    if not any(lo <= nargs <= hi for lo, hi in want):
                        bad.append(f"{os.path.relpath(path, ROOT)}:{line}: mod.{fn}() passes {nargs} args, bindings take {sorted(want)}")
    assert not bad, "\n".join(bad)
    seen = {fn for path in files for recv, fn, _, _ in _calls(path) if recv in RECEIVERS['cuda'] | {'mod', 'self.mod'}}
    assert {'twoshot_fedavg', 'gemm_tcgen05', 'produced_mark', 'star_round'} <= seen, seen
    assert checked >= 40, checked                                # the scan really found the call sites


def test_extension_accepts_the_argument_types_the_python_layer_passes():
    """pybind converts the arguments before the function body runs: on a CPU-only box the body then fails (no device, CPU
    tensors) with a RuntimeError — a TypeError would mean the Python layer passes something the binding cannot take."""
    import torch

    from colearn_federated_learning_b200.ops import _ext
    ext = _ext.load()
    if ext is None:
        pytest.skip("CUDA extension not built")
    if torch.cuda.is_available():
        pytest.skip("meant for the CPU-only box (the calls below would launch kernels on garbage pointers)")
    ptrs = [4096, 8192]
    # parallel/engine.py::_run_twoshot (classic and overlapped form)
    for produced, timeout, arrive in ((0, 0.0, [16, 32]), (123456, 20.0, [])):
        with pytest.raises(RuntimeError):
            ext.twoshot_fedavg(ptrs, ptrs, ptrs, 64, 128, 0, 3, 0b11, 1.0, 4096, 2048, 0, 16, arrive, True, 0, 0, produced, timeout)
    with pytest.raises(RuntimeError):
        ext.produced_mark(4096, 2048, 0, 100)
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    m = torch.zeros(128, 128)
    # ops/linear.py::gemm_bf16 with every optional group in the form it is passed (CPU tensors are rejected by a TORCH_CHECK)
    for prod in ([], [0, 0, 132], [4096, 512, 132], [0, 0, 0, 1], [4096, 512, 132, 0]):
        with pytest.raises(RuntimeError):
            ext.gemm_tcgen05(a, a, None, False, None, None, None, None, m, 0.1, None, None, None, 0, 0, 1, 0, 0, 0, 0, 0, None, 0, False, None, [], prod)
    packed = ext.produced_signal_pack(4096, [8192, 12288], 16384, 1, 2048, 1, 3, 6000)
    assert packed.dtype == torch.uint8 and packed.numel() >= 160
