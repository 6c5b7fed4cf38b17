"""Plain-PyTorch fp32 definitions of every op in ``ops/`` .

Two jobs: (1) the numerics oracle each sm_100a kernel is tested against, (2) the CPU execution
path (BASELINE config 1 runs with no GPU; multi-process CPU tests run over gloo).  Nothing here
is on the GPU hot path — on a CUDA box the dispatchers in ``ops/__init__`` refuse to fall back
to these silently.

The local-SGD semantics re-create PySyft's ``FederatedClient._fit`` ([EXTERNAL]):
``for epoch: for batch in shuffled(data): zero_grad; out = model(x); loss = loss_fn(out, y);
backward; SGD step; if ++it >= max_nr_batches >= 0: stop`` and return the last loss.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

LOSSES = ("bce", "sse", "xent", "mse")


# ------------------------------------------------------------------------------------------------
# flat-arena helpers
# ------------------------------------------------------------------------------------------------
def mlp_views(flat: torch.Tensor, dims: Sequence[int]) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """[(W[out,in], b[out])] views into the flat arena (state-dict order)."""
    out, off = [], 0
    for i in range(len(dims) - 1):
        k, n = dims[i], dims[i + 1]
        w = flat[off:off + k * n].view(n, k)
        off += k * n
        b = flat[off:off + n]
        off += n
        out.append((w, b))
    return out


def mlp_forward(flat: torch.Tensor, dims: Sequence[int], x: torch.Tensor,
                out_activation: str = "none") -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Returns (output, [input of every layer]) — activations are kept for the backward."""
    acts = []
    h = x
    views = mlp_views(flat, dims)
    for li, (w, b) in enumerate(views):
        acts.append(h)
        h = h @ w.t() + b
        if li < len(views) - 1:
            h = torch.relu(h)
    if out_activation == "sigmoid":
        h = torch.sigmoid(h)
    return h, acts


def loss_and_dz(out: torch.Tensor, y: torch.Tensor, loss: str, out_activation: str
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Loss value and gradient w.r.t. the last layer's *pre-activation*.

    bce  : mean binary cross-entropy on probabilities (reference remote mode, cf.py:76-79).
           Needs a sigmoid head; log terms are clamped at -100 like ``F.binary_cross_entropy``.
    sse  : sum of squared errors (reference local mode, cf.py:112).
    mse  : sse / batch (reference encrypted mode, cf.py:158).
    xent : softmax cross-entropy, integer labels in ``y[:, 0]``, mean over the batch.
    """
    bsz = out.shape[0]
    if loss == "xent":
        labels = y.reshape(bsz).long()
        lse = torch.logsumexp(out, dim=1)
        value = (lse - out.gather(1, labels.view(-1, 1)).squeeze(1)).mean()
        p = torch.softmax(out, dim=1)
        p = p - torch.nn.functional.one_hot(labels, out.shape[1]).to(p.dtype)
        return value, p / bsz
    y = y.reshape(out.shape).to(out.dtype)
    if loss == "bce":
        if out_activation != "sigmoid":
            raise ValueError("bce needs a sigmoid head")
        logp = torch.clamp(torch.log(out), min=-100.0)
        log1mp = torch.clamp(torch.log1p(-out), min=-100.0)
        value = -(y * logp + (1 - y) * log1mp).mean()
        return value, (out - y) / out.numel()
    if loss in ("sse", "mse"):
        diff = out - y
        scale = 1.0 if loss == "sse" else 1.0 / bsz
        value = (diff * diff).sum() * scale
        dz = 2.0 * diff * scale
        if out_activation == "sigmoid":
            dz = dz * out * (1 - out)
        return value, dz
    raise ValueError(f"unknown loss {loss!r}")


def mlp_backward(flat: torch.Tensor, dims: Sequence[int], acts: List[torch.Tensor],
                 dz: torch.Tensor) -> torch.Tensor:
    """Gradient of the flat arena given d(loss)/d(pre-activation of the last layer)."""
    grad = torch.zeros_like(flat)
    gviews = mlp_views(grad, dims)
    views = mlp_views(flat, dims)
    for li in range(len(views) - 1, -1, -1):
        w, _ = views[li]
        gw, gb = gviews[li]
        a = acts[li]
        gw.copy_(dz.t() @ a)
        gb.copy_(dz.sum(0))
        if li > 0:
            dz = (dz @ w) * (a > 0).to(dz.dtype)  # a = relu(z_{li-1})
    return grad


def mlp_grad(flat, dims, x, y, loss="xent", out_activation="none"):
    out, acts = mlp_forward(flat, dims, x, out_activation)
    value, dz = loss_and_dz(out, y, loss, out_activation)
    return value, mlp_backward(flat, dims, acts, dz)


# ------------------------------------------------------------------------------------------------
# local SGD (the worker-side fit loop)
# ------------------------------------------------------------------------------------------------
def make_permutation(n: int, epochs: int, seed: int, device=None, shuffle: bool = True) -> torch.Tensor:
    """int32 ``[epochs, n]`` sample order.  CPU generator => identical on every backend."""
    g = torch.Generator().manual_seed(int(seed))
    rows = [torch.randperm(n, generator=g) if shuffle else torch.arange(n) for _ in range(epochs)]
    perm = torch.stack(rows).to(torch.int32)
    return perm.to(device) if device is not None else perm


def total_steps(n: int, batch_size: int, epochs: int, max_nr_batches: int = -1) -> int:
    """Number of SGD steps a fit performs (``max_nr_batches <= 0`` means unlimited)."""
    steps = epochs * math.ceil(n / batch_size)
    if max_nr_batches is not None and max_nr_batches > 0:
        steps = min(steps, max_nr_batches)
    return steps


def mlp_local_sgd(flat: torch.Tensor, dims: Sequence[int], x: torch.Tensor, y: torch.Tensor,
                  perm: torch.Tensor, batch_size: int = 1, lr: float = 0.01, epochs: int = 1,
                  max_nr_batches: int = -1, loss: str = "xent", out_activation: str = "none"
                  ) -> torch.Tensor:
    """In-place local SGD on the flat arena; returns the last batch's loss (0-dim tensor).

    ``max_nr_batches``: PySyft stops when ``it >= max_nr_batches >= 0`` *after* a step, so 0
    behaves like "one step"; negative means unlimited.  We normalise: <= 0 → unlimited except
    that the reference never passes 0 (default -1, round mode 1000)."""
    n = x.shape[0]
    it = 0
    last = torch.zeros((), dtype=flat.dtype, device=flat.device)
    limit = max_nr_batches if (max_nr_batches is not None and max_nr_batches > 0) else None
    for e in range(epochs):
        order = perm[e % perm.shape[0]].long()
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            last, g = mlp_grad(flat, dims, x[idx], y[idx], loss, out_activation)
            flat.sub_(lr * g)
            it += 1
            if limit is not None and it >= limit:
                return last
    return last


# ------------------------------------------------------------------------------------------------
# standalone fused-op references
# ------------------------------------------------------------------------------------------------
def sigmoid_bce(z: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean BCE of sigmoid(z), d/dz)."""
    p = torch.sigmoid(z)
    return loss_and_dz(p, y, "bce", "sigmoid")


def softmax_xent(logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return loss_and_dz(logits.float(), labels.view(-1, 1), "xent", "none")


def sse_loss(out: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return loss_and_dz(out, y, "sse", "none")


def sgd_step(param: torch.Tensor, grad: torch.Tensor, lr: float) -> torch.Tensor:
    return param.sub_(grad.to(param.dtype), alpha=lr)


def fedavg_flat(models: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """``sum_k w_k * models[k]``; models ``[K, P]``, weights ``[K]`` (already normalised)."""
    return (weights.view(-1, 1).to(models.dtype) * models).sum(0)


def fedavg_apply(theta: torch.Tensor, models: torch.Tensor, weights: torch.Tensor,
                 server_lr: float = 1.0) -> torch.Tensor:
    """``theta <- theta + server_lr * (sum_k w_k models[k] - theta)`` (weights sum to 1)."""
    avg = fedavg_flat(models, weights)
    return theta.add_(server_lr * (avg - theta))


def linear_bias_relu(x, w, b, relu=True):
    y = x.float() @ w.float().t() + b.float()
    return torch.relu(y) if relu else y


def eval_binary(out: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(sum of per-sample BCE, number of correct ``round(out)`` predictions) — cf.py:233-253."""
    y = y.reshape(out.shape).to(out.dtype)
    logp = torch.clamp(torch.log(out), min=-100.0)
    log1mp = torch.clamp(torch.log1p(-out), min=-100.0)
    loss = -(y * logp + (1 - y) * log1mp).sum()
    correct = (torch.round(out) == y).sum()
    return loss, correct


def argmax_rows(out: torch.Tensor) -> torch.Tensor:
    return out.argmax(1, keepdim=True)
