"""Native CPU executor of the MLP local fit (ops/csrc/mlp_host.cpp) against the PyTorch definitions (ops/reference.py)."""
import os
import time

import pytest
import torch

from colearn_federated_learning_b200 import ops
from colearn_federated_learning_b200.models import build_model, flatten_params
from colearn_federated_learning_b200.ops import host, reference


@pytest.fixture(scope="module", autouse=True)
def _built():
    if host.load(build_if_missing=True) is None:
        pytest.skip(f"host executor could not be built here: {host.last_error()}")


def _data(dims, loss, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, dims[0], generator=g)
    if loss == "xent":
        y = torch.randint(0, dims[-1], (n, 1), generator=g).float()
    else:
        y = (torch.rand(n, dims[-1], generator=g) > 0.5).float()
    return x, y


CASES = [
    ((10, 50, 30, 10, 1), "sigmoid", "bce"),      # reference FFNN, remote mode
    ((10, 50, 30, 10, 1), "sigmoid", "sse"),      # reference FFNN, local mode
    ((10, 50, 30, 10, 1), "sigmoid", "mse"),      # encrypted-mode loss
    ((10, 64, 64, 2), "none", "xent"),            # BASELINE configs 2/3
    ((2, 50, 10, 1), "none", "sse"),              # TestingRemote / XOR
    ((10, 7, 3), "none", "xent"),                 # an architecture with no GPU instantiation
]


@pytest.mark.parametrize("dims,act,loss", CASES)
@pytest.mark.parametrize("batch", [1, 4, 7])
def test_host_fit_matches_reference(dims, act, loss, batch):
    torch.manual_seed(1)
    n = 37
    x, y = _data(dims, loss, n)
    p = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
    flat0 = torch.randn(p) * 0.3
    perm = reference.make_permutation(n, 2, seed=5)
    want = flat0.clone()
    last_ref = reference.mlp_local_sgd(want, dims, x, y, perm, batch, 0.05, 3, -1, loss, act)     # 3 epochs over 2 perm rows
    got = flat0.clone()
    res = host.mlp_local_sgd_multi(dims, act, [got], [x], [y], [perm], batch, 0.05, 3, -1, loss)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)
    assert abs(float(res[0, 0]) - float(last_ref)) < 1e-4 * max(1.0, abs(float(last_ref)))


def test_max_nr_batches_identity_order_and_mean_loss():
    dims, act, loss = (10, 64, 64, 2), "none", "xent"
    x, y = _data(dims, loss, 50)
    flat0 = flatten_params(build_model("mlp")).clone()
    want, got = flat0.clone(), flat0.clone()
    ident = reference.make_permutation(50, 1, 0, shuffle=False)
    losses = []
    w2 = flat0.clone()
    for i in range(9):                                       # the mean the executor reports = mean of the step losses
        v, g = reference.mlp_grad(w2, dims, x[i:i + 1], y[i:i + 1], loss, act)
        w2.sub_(0.1 * g)
        losses.append(float(v))
    reference.mlp_local_sgd(want, dims, x, y, ident, 1, 0.1, 5, 9, loss, act)
    res = host.mlp_local_sgd_multi(dims, act, [got], [x], [y], [None], 1, 0.1, 5, 9, loss)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got, w2, rtol=1e-4, atol=1e-5)
    assert abs(float(res[0, 1]) - sum(losses) / 9) < 1e-4 and abs(float(res[0, 0]) - losses[-1]) < 1e-4


def test_clients_train_concurrently_and_independently():
    dims, act, loss = (10, 50, 30, 10, 1), "sigmoid", "bce"
    k, n = 6, 64
    flat0 = flatten_params(build_model("ffnn")).clone()
    flats = flat0.unsqueeze(0).repeat(k, 1).contiguous()
    xs, ys, perms = [], [], []
    for i in range(k):
        x, y = _data(dims, loss, n + i, seed=i)
        xs.append(x), ys.append(y), perms.append(reference.make_permutation(n + i, 1, seed=i))
    res = host.mlp_local_sgd_multi(dims, act, [flats[i] for i in range(k)], xs, ys, perms, 1, 0.01, 1, -1, loss, threads=3)
    for i in range(k):
        want = flat0.clone()
        last = reference.mlp_local_sgd(want, dims, xs[i], ys[i], perms[i], 1, 0.01, 1, -1, loss, act)
        torch.testing.assert_close(flats[i], want, rtol=2e-4, atol=2e-5)
        assert abs(float(res[i, 0]) - float(last)) < 1e-4


def test_dispatch_and_opt_out(monkeypatch):
    """``ops.mlp_local_sgd`` on CPU tensors runs the executor; ``COLEARN_HOST_KERNELS=0`` keeps the definitions."""
    dims = (10, 64, 64, 2)
    x, y = _data(dims, "xent", 40)
    flat0 = flatten_params(build_model("mlp")).clone()
    perm = reference.make_permutation(40, 1, 3)
    a, b = flat0.clone(), flat0.clone()
    la = ops.mlp_local_sgd(a, dims, x, y, perm, 1, 0.05, 1, -1, "xent", "none")
    monkeypatch.setenv("COLEARN_HOST_KERNELS", "0")
    lb = ops.mlp_local_sgd(b, dims, x, y, perm, 1, 0.05, 1, -1, "xent", "none")
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    assert abs(float(la) - float(lb)) < 1e-4 and not torch.equal(a, flat0)


def test_forward_and_argument_checks():
    dims = (10, 50, 30, 10, 1)
    flat = flatten_params(build_model("ffnn")).clone()
    x, y = _data(dims, "bce", 9)
    torch.testing.assert_close(host.mlp_forward(flat, dims, x, "sigmoid"), reference.mlp_forward(flat, dims, x, "sigmoid")[0],
                               rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        host.mlp_local_sgd_multi(dims, "none", [flat.clone()], [x], [y], [None], loss="bce")          # bce needs a sigmoid head
    with pytest.raises(RuntimeError):
        host.mlp_local_sgd_multi(dims, "sigmoid", [flat[:-1].clone()], [x], [y], [None], loss="bce")  # wrong arena size
    with pytest.raises(RuntimeError):
        bad = torch.full((1, 9), 9, dtype=torch.int32)
        host.mlp_local_sgd_multi(dims, "sigmoid", [flat.clone()], [x], [y], [bad], loss="bce")        # perm out of range
    with pytest.raises(ValueError):
        host.mlp_local_sgd_multi(dims, "sigmoid", [flat.double()], [x], [y], [None], loss="bce")


def test_host_executor_is_orders_of_magnitude_faster_than_eager_definitions():
    dims = (10, 64, 64, 2)
    x, y = _data(dims, "xent", 400)
    flat = flatten_params(build_model("mlp")).clone()
    t0 = time.perf_counter()
    host.mlp_local_sgd_multi(dims, "none", [flat.clone()], [x], [y], [None], 1, 0.01, 1, -1, "xent")
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    reference.mlp_local_sgd(flat.clone(), dims, x, y, reference.make_permutation(400, 1, 0, shuffle=False), 1, 0.01, 1, -1, "xent")
    t_ref = time.perf_counter() - t0
    assert t_host * 5 < t_ref, (t_host, t_ref)


def test_host_fit_matches_reference_on_random_architectures():
    """Property test: any MLP shape / batch size / loss / step limit — the executor tracks the PyTorch definitions."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(hidden=st.lists(st.integers(1, 24), min_size=0, max_size=3), d_in=st.integers(1, 12), d_out=st.integers(1, 5),
           batch=st.integers(1, 9), n=st.integers(1, 23), loss=st.sampled_from(["xent", "sse", "mse", "bce"]),
           epochs=st.integers(1, 3), limit=st.sampled_from([-1, 1, 4]), seed=st.integers(0, 10_000))
    def prop(hidden, d_in, d_out, batch, n, loss, epochs, limit, seed):
        dims = (d_in, *hidden, d_out)
        act = "sigmoid" if loss == "bce" else ("sigmoid" if seed % 3 == 0 and loss != "xent" else "none")
        if loss == "xent" and d_out < 2:
            dims = (d_in, *hidden, 2)
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(n, dims[0], generator=g)
        y = (torch.randint(0, dims[-1], (n, 1), generator=g).float() if loss == "xent"
             else (torch.rand(n, dims[-1], generator=g) > 0.5).float())
        p = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
        flat0 = (torch.rand(p, generator=g) - 0.5) * 0.8
        perm = reference.make_permutation(n, epochs, seed)
        want, got = flat0.clone(), flat0.clone()
        last = reference.mlp_local_sgd(want, dims, x, y, perm, batch, 0.05, epochs, limit, loss, act)
        res = host.mlp_local_sgd_multi(dims, act, [got], [x], [y], [perm], batch, 0.05, epochs, limit, loss)
        torch.testing.assert_close(got, want, rtol=5e-4, atol=5e-5)
        assert abs(float(res[0, 0]) - float(last)) <= 2e-4 * max(1.0, abs(float(last)))

    prop()


def test_fixed_shape_fast_path_agrees_with_the_generic_loop():
    """The reference's three networks take a compile-time-shape path at batch 1 (csrc/mlp_host.cpp: run_fit_fixed); the
    generic loop (COLEARN_HOST_GENERIC=1, read once per process) must agree with it to rounding."""
    import subprocess
    import sys
    code = (
        "import torch, hashlib\n"
        "from colearn_federated_learning_b200.ops import host, reference as R\n"
        "from colearn_federated_learning_b200.models import MLP, FFNN, TestingRemote, flatten_params\n"
        "torch.set_num_threads(1)\n"
        "for ctor, loss in ((MLP, 'xent'), (FFNN, 'bce'), (TestingRemote, 'mse')):\n"
        "    torch.manual_seed(3); m = ctor()\n"
        "    dims, act = m.spec.dims, m.spec.out_activation\n"
        "    x = torch.rand(200, dims[0]); y = (torch.rand(200, 1) > 0.5).float()\n"
        "    flat = flatten_params(m).clone()\n"
        "    out = host.mlp_local_sgd_multi(dims, act, [flat], [x], [y], [R.make_permutation(200, 1, 0)], 1, 0.05, 1, -1, loss)\n"
        "    print('RES', ctor.__name__, ' '.join(f'{v:.9e}' for v in flat[:64].tolist()), f'{float(out[0, 0]):.9e}')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for generic in (False, True):
        env = dict(os.environ)
        env.pop("COLEARN_HOST_GENERIC", None)
        if generic:
            env["COLEARN_HOST_GENERIC"] = "1"
        p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-1500:]
        outs.append([ln.split()[2:] for ln in p.stdout.splitlines() if ln.startswith("RES")])
    assert len(outs[0]) == len(outs[1]) == 3
    for a, b in zip(outs[0], outs[1]):
        va, vb = torch.tensor([float(v) for v in a]), torch.tensor([float(v) for v in b])
        assert torch.allclose(va, vb, rtol=1e-4, atol=1e-5), float((va - vb).abs().max())
