"""Model registry + flat parameter arena (un)flatten.

All model state that FedAvg touches lives in ONE contiguous fp32 vector ("arena") in
state-dict order, so "send model" and "average models" are single contiguous messages
(SURVEY §7.0).  ``named_parameters()`` only — buffers are not averaged (reference
``utils.federated_avg`` semantics).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .mlp import FFNN, MLP, MLPNet, MLPSpec, Net, TestingRemote, WideMLP
from .resnet import ResNet18

MODEL_REGISTRY: Dict[str, Callable[[], nn.Module]] = {
    "ffnn": FFNN,
    "testing_remote": TestingRemote,
    "net": Net,
    "mlp": MLP,
    "wide_mlp": WideMLP,
    "resnet18": ResNet18,
}

DEFAULT_LOSS = {"ffnn": "bce", "testing_remote": "sse", "net": "xent", "mlp": "xent",
                "wide_mlp": "xent", "resnet18": "xent"}


def build_model(name: str, **kwargs) -> nn.Module:
    try:
        return MODEL_REGISTRY[name](**kwargs)
    except KeyError:
        raise ValueError(f"unknown model {name!r}; choose from {sorted(MODEL_REGISTRY)}") from None


def model_spec(model: nn.Module) -> Optional[MLPSpec]:
    return model.spec if isinstance(model, MLPNet) else None


def param_layout(model: nn.Module) -> List[Tuple[str, torch.Size, int, int]]:
    """[(name, shape, offset, numel)] for ``named_parameters()`` in order."""
    out, off = [], 0
    for name, p in model.named_parameters():
        out.append((name, p.shape, off, p.numel()))
        off += p.numel()
    return out


def num_params(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


@torch.no_grad()
def flatten_params(model: nn.Module, out: Optional[torch.Tensor] = None,
                   dtype: torch.dtype = torch.float32) -> torch.Tensor:
    n = num_params(model)
    if out is None:
        dev = next(model.parameters()).device
        out = torch.empty(n, dtype=dtype, device=dev)
    off = 0
    for p in model.parameters():
        k = p.numel()
        out[off:off + k].copy_(p.detach().reshape(-1))
        off += k
    return out


@torch.no_grad()
def unflatten_params(model: nn.Module, flat: torch.Tensor) -> nn.Module:
    off = 0
    for p in model.parameters():
        k = p.numel()
        p.copy_(flat[off:off + k].view_as(p))
        off += k
    return model


def alias_params_to_arena(model: nn.Module, flat: torch.Tensor) -> nn.Module:
    """Re-point every parameter's storage into ``flat`` (zero-copy).  After this, the kernels
    that update ``flat`` (broadcast, SGD, FedAvg apply) update the module in place — the
    arena-era equivalent of PySyft's in-place ``model.send()/get()``."""
    off = 0
    with torch.no_grad():
        for p in model.parameters():
            k = p.numel()
            view = flat[off:off + k].view(p.shape)
            view.copy_(p.detach())
            p.data = view
            off += k
    return model


def state_dict_from_flat(model: nn.Module, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
    """CPU fp32 state_dict (params from ``flat``, buffers from the module)."""
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    off = 0
    for name, p in model.named_parameters():
        k = p.numel()
        sd[name] = flat[off:off + k].detach().float().cpu().view(p.shape).clone()
        off += k
    return sd
