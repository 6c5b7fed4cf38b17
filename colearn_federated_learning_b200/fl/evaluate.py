"""Evaluation and inference helpers.

``evaluate`` mirrors the reference helper (``client_federated.py:217-253``: BCE sum / N and
``round(out)`` accuracy; its only call site is commented out at fc.py:587-590) without the
batch-size-1 assumption and without printing every mis-prediction.  ``predict`` is the on-worker
inference of fc.py:252-256 (forward + ``argmax(1, keepdim=True)``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import ops
from ..models import MLPNet, flatten_params


@torch.no_grad()
def forward_flat(model: nn.Module, flat: Optional[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    if isinstance(model, MLPNet) and flat is not None:
        spec = model.spec
        xx = x.view(x.shape[0], -1) if spec.flatten_input else x
        if not flat.is_cuda or ops.net_kind_for(spec.dims, spec.out_activation) is not None:
            return ops.mlp_forward(flat, spec.dims, xx.to(flat.device), spec.out_activation)
    if flat is not None and flat.is_cuda and x.dim() == 4:
        # conv nets on a GPU: eval-mode forward on this repo's kernels (fl/convnet.py), batch 128 chunks
        from .convnet import ConvNetTrainer
        from .trainer import conv_path
        from ..models.resnet import ResNet18
        if isinstance(model, ResNet18) and conv_path(flat.device) == "native" and x.shape[2] >= 32 and x.shape[3] >= 32:
            tr = ConvNetTrainer.cached(model, flat, 128, tuple(x.shape[2:]))
            tr.load(flat, model)
            return tr.infer(x.to(flat.device))
    model.eval()
    return model(x)


@torch.no_grad()
def evaluate(model: nn.Module, x: torch.Tensor, y: torch.Tensor, flat: Optional[torch.Tensor] = None,
             verbose: bool = True) -> Dict[str, float]:
    """Binary models: average BCE + accuracy of ``round(out)``; multi-class: xent + argmax accuracy."""
    if flat is None and isinstance(model, MLPNet):
        flat = flatten_params(model).to(x.device)
    out = forward_flat(model, flat, x)
    n = x.shape[0]
    if out.shape[1] == 1 and getattr(getattr(model, "spec", None), "out_activation", "") == "sigmoid":
        loss_sum, correct = ops.eval_binary(out.reshape(-1), y.to(out.device).reshape(-1).float())
        res = {"loss": float(loss_sum) / n, "correct": int(correct), "n": n, "accuracy": int(correct) / n}
    else:
        labels = y.to(out.device).reshape(-1).long()
        loss, _ = ops.softmax_xent(out.float().contiguous(), labels)
        correct = int((ops.argmax_rows(out.float().contiguous()).reshape(-1) == labels).sum())
        res = {"loss": float(loss), "correct": correct, "n": n, "accuracy": correct / n}
    if verbose:
        print("\nTest set: Average loss: {:.4f}, Accuracy: {}/{} ({:.5f}%)\n".format(
            res["loss"], res["correct"], n, 100.0 * res["accuracy"]))
    return res


@torch.no_grad()
def predict(model: nn.Module, x: torch.Tensor, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``model(x).argmax(1, keepdim=True)`` — what the coordinator logs for an INFERENCE event."""
    out = forward_flat(model, flat, x)
    return ops.argmax_rows(out.float().contiguous())
