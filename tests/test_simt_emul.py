"""The persistent-MLP CUDA kernels (csrc/mlp_persistent.cu + mlp_v2.inc — the flagship op of BASELINE configs 1-3)
executed on the CPU: the kernel *source* is compiled with g++ through csrc/host_shim.h (one OS thread per CUDA thread,
__syncthreads / warp shuffles / dynamic shared memory emulated) and compared with the PyTorch definitions, exactly
like tests/test_gpu_kernels.py does on a B200.  Every variant: v1 (weights in shared memory, 256 threads), v2 with 256
and with 128 threads (register-resident row + column copies, replicated head)."""
import pytest
import torch

from colearn_federated_learning_b200.models import build_model, flatten_params
from colearn_federated_learning_b200.ops import reference

NETS = {  # kind: (model name, dims, output activation, losses)
    0: ("ffnn", (10, 50, 30, 10, 1), "sigmoid", ("bce", "sse", "mse")),
    1: ("mlp", (10, 64, 64, 2), "none", ("xent",)),
    2: ("testing_remote", (2, 50, 10, 1), "none", ("sse", "mse")),
}
LOSS_CODES = {"bce": 0, "sse": 1, "xent": 2, "mse": 3}


@pytest.fixture(scope="module")
def simt():
    import glob
    import importlib.util
    import os

    from colearn_federated_learning_b200.ops import build
    try:
        path = build.build_simt_emul()
    except Exception as e:  # noqa: BLE001 - no C++20 compiler / CUDA headers on this box
        pytest.skip(f"SIMT emulator could not be built here: {e}")
    spec = importlib.util.spec_from_file_location("_colearn_simt", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert glob.glob(os.path.join(os.path.dirname(path), "_colearn_simt*.so"))
    return mod


def _data(dims, loss, n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, dims[0], generator=g)
    y = (torch.randint(0, dims[-1], (n, 1), generator=g).float() if loss == "xent"
         else (torch.rand(n, dims[-1], generator=g) > 0.5).float())
    return x, y


CASES = [(k, loss, v, b) for k, (_, _, _, losses) in NETS.items() for loss in losses for v, b in ((3, 1), (2, 1), (1, 1), (3, 4))]


@pytest.mark.parametrize("kind,loss,variant,batch", CASES)
def test_kernel_source_on_cpu_matches_reference(simt, kind, loss, variant, batch):
    name, dims, act, _ = NETS[kind]
    n = 23
    x, y = _data(dims, loss, n, seed=kind * 10 + batch)
    torch.manual_seed(kind)
    flat0 = flatten_params(build_model(name)).clone()
    assert simt.mlp_net_params(kind) == flat0.numel()
    perm = reference.make_permutation(n, 2, seed=5)
    want = flat0.clone()
    last = reference.mlp_local_sgd(want, dims, x, y, perm, batch, 0.05, 2, -1, loss, act)
    got = flat0.clone()
    losses = simt.mlp_local_sgd(kind, flat0, [got], [x], [y], [perm], batch, 2, -1, LOSS_CODES[loss], 0.05, variant, [1.0], False, None)
    torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-4)
    assert abs(float(losses[0, 0]) - float(last)) < 2e-3 * max(1.0, abs(float(last)))


def test_step_limit_scale_delta_flags_and_many_clients(simt):
    """max_nr_batches, the FedAvg weight pre-applied by the producer (out = w * theta_k or w * (theta_k - theta_in)), the
    completion flag (st.release) and one CTA per client — the star protocol's worker side."""
    kind, (name, dims, act, _) = 1, NETS[1]
    flat0 = flatten_params(build_model(name)).clone()
    k = 3
    xs, ys, perms, wants = [], [], [], []
    for i in range(k):
        x, y = _data(dims, "xent", 17 + i, seed=i)
        p = reference.make_permutation(17 + i, 1, seed=i)
        w = flat0.clone()
        reference.mlp_local_sgd(w, dims, x, y, p, 1, 0.1, 3, 11, "xent", act)       # 3 epochs, capped at 11 steps
        xs.append(x), ys.append(y), perms.append(p), wants.append(w)
    outs = [torch.zeros_like(flat0) for _ in range(k)]
    flags = torch.zeros(k, dtype=torch.int32)
    scales = [0.5, 0.25, 0.25]
    simt.mlp_local_sgd(kind, flat0, outs, xs, ys, perms, 1, 3, 11, LOSS_CODES["xent"], 0.1, 3, scales, False, flags)
    for i in range(k):
        torch.testing.assert_close(outs[i], scales[i] * wants[i], rtol=2e-3, atol=2e-4)
    assert flags.tolist() == [7, 8, 9]
    deltas = [torch.zeros_like(flat0) for _ in range(k)]
    simt.mlp_local_sgd(kind, flat0, deltas, xs, ys, perms, 1, 3, 11, LOSS_CODES["xent"], 0.1, 3, scales, True, None)
    for i in range(k):
        torch.testing.assert_close(deltas[i], scales[i] * (wants[i] - flat0), rtol=2e-3, atol=2e-4)
    # FedAvg of the pre-scaled contributions = the weighted mean of the locally trained models
    torch.testing.assert_close(sum(outs), sum(s * w for s, w in zip(scales, wants)), rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_forward_kernel_on_cpu(simt, kind):
    name, dims, act, _ = NETS[kind]
    flat = flatten_params(build_model(name)).clone()
    x = torch.rand(41, dims[0])
    got = simt.mlp_forward(kind, flat, x, dims[-1])
    torch.testing.assert_close(got, reference.mlp_forward(flat, dims, x, act)[0], rtol=1e-4, atol=1e-5)
