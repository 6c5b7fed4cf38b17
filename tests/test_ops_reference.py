"""The pure-PyTorch op definitions (numerics oracle + CPU path) against torch autograd."""
import pytest
import torch
import torch.nn.functional as F

from colearn_federated_learning_b200.models import FFNN, MLP, TestingRemote, flatten_params
from colearn_federated_learning_b200.ops import reference as R


@pytest.mark.parametrize("ctor,loss", [(FFNN, "bce"), (FFNN, "sse"), (MLP, "xent"), (TestingRemote, "sse"), (TestingRemote, "mse")])
@pytest.mark.parametrize("bsz", [1, 7])
def test_manual_grad_matches_autograd(ctor, loss, bsz):
    torch.manual_seed(0)
    model = ctor()
    spec = model.spec
    flat = flatten_params(model).clone()
    x = torch.rand(bsz, spec.dims[0])
    if loss == "xent":
        y = torch.randint(0, spec.dims[-1], (bsz, 1)).float()
    else:
        y = (torch.rand(bsz, spec.dims[-1]) > 0.5).float()
    value, grad = R.mlp_grad(flat, spec.dims, x, y, loss, spec.out_activation)

    out = model(x)
    if loss == "bce":
        ref = F.binary_cross_entropy(out, y)
    elif loss == "sse":
        ref = ((out - y) ** 2).sum()
    elif loss == "mse":
        ref = ((out - y) ** 2).sum() / bsz
    else:
        ref = F.cross_entropy(out, y.view(-1).long())
    ref.backward()
    ref_grad = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(value, ref.detach(), atol=1e-6, rtol=1e-5)
    assert torch.allclose(grad, ref_grad, atol=1e-6, rtol=1e-4)


def test_local_sgd_matches_torch_optimizer():
    torch.manual_seed(1)
    model = FFNN()
    flat = flatten_params(model).clone()
    x, y = torch.rand(23, 10), (torch.rand(23, 1) > 0.5).float()
    perm = R.make_permutation(23, 2, seed=3)
    last = R.mlp_local_sgd(flat, model.spec.dims, x, y, perm, batch_size=4, lr=0.05, epochs=2, loss="bce",
                           out_activation="sigmoid")
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    for e in range(2):
        order = perm[e].long()
        for lo in range(0, 23, 4):
            idx = order[lo:lo + 4]
            opt.zero_grad()
            loss = F.binary_cross_entropy(model(x[idx]), y[idx])
            loss.backward()
            opt.step()
    assert torch.allclose(flat, flatten_params(model), atol=1e-6, rtol=1e-4)
    assert torch.allclose(last, loss.detach(), atol=1e-6)


def test_max_nr_batches_limits_steps():
    torch.manual_seed(2)
    model = MLP()
    x, y = torch.rand(40, 10), torch.randint(0, 2, (40, 1)).float()
    perm = R.make_permutation(40, 1, 0)
    a, b = flatten_params(model).clone(), flatten_params(model).clone()
    R.mlp_local_sgd(a, model.spec.dims, x, y, perm, 1, 0.1, 1, max_nr_batches=5)
    R.mlp_local_sgd(b, model.spec.dims, x[perm[0, :5].long()], y[perm[0, :5].long()],
                    R.make_permutation(5, 1, 0, shuffle=False), 1, 0.1, 1, -1)
    assert torch.allclose(a, b)
    assert R.total_steps(40, 1, 1, 5) == 5 and R.total_steps(40, 8, 2, -1) == 10


def test_fedavg_apply_reference():
    theta = torch.zeros(8)
    models = torch.stack([torch.full((8,), 1.0), torch.full((8,), 3.0)])
    R.fedavg_apply(theta, models, torch.tensor([0.25, 0.75]), 1.0)
    assert torch.allclose(theta, torch.full((8,), 2.5))
