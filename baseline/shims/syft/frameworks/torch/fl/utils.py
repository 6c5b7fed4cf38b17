"""``syft.frameworks.torch.fl.utils`` stand-in: unweighted parameter mean, accumulated into the FIRST model."""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch import nn


def add_model(dst_model: nn.Module, src_model: nn.Module) -> nn.Module:
    """``dst += src`` over the parameters both models name (buffers are not touched)."""
    dst = dict(dst_model.named_parameters())
    with torch.no_grad():
        for name, p in src_model.named_parameters():
            if name in dst:
                dst[name].set_(p.data + dst[name].data)
    return dst_model


def scale_model(model: nn.Module, scale: float) -> nn.Module:
    with torch.no_grad():
        for p in model.parameters():
            p.set_(p.data * scale)
    return model


def federated_avg(models: Dict[Any, nn.Module]) -> nn.Module:
    """Mean of the models' parameters with uniform weights, in place in ``models``' first entry."""
    model_list = list(models.values())
    model = model_list[0]
    for other in model_list[1:]:
        model = add_model(model, other)
    return scale_model(model, 1.0 / len(model_list))
