"""Federated averaging.

Parity: ``utils.federated_avg(models: dict id -> module)`` (call sites fc.py:373,568; PySyft
0.2.x semantics recalled in SURVEY C16 / §2.3): the **unweighted** arithmetic mean of
``named_parameters()`` accumulated in place into the first model; buffers are not averaged.
``BASELINE.json`` asks for the sample-count-weighted generalisation ``w_k = n_k / sum(n)``;
uniform weights reproduce the reference exactly (tested).

The module-level API below works on ``nn.Module`` dicts for CLI/CPU use; the data-plane version
(``ops.fedavg_apply`` / ``parallel.collectives.fedavg_reduce_apply``) works on flat arenas.
"""
from __future__ import annotations

from typing import Mapping, Optional, Sequence

import torch
import torch.nn as nn


def normalized_weights(sample_counts: Optional[Sequence[float]], k: int,
                       device=None, dtype=torch.float32) -> torch.Tensor:
    """``n_k / sum(n)``; ``None`` → uniform ``1/K`` (reference)."""
    if sample_counts is None:
        return torch.full((k,), 1.0 / k, dtype=dtype, device=device)
    w = torch.as_tensor(list(sample_counts), dtype=torch.float64)
    if len(w) != k:
        raise ValueError("one sample count per model required")
    s = w.sum()
    if s <= 0:
        raise ValueError("sample counts must sum to a positive number")
    return (w / s).to(dtype=dtype, device=device)


@torch.no_grad()
def federated_avg(models: Mapping[str, nn.Module],
                  sample_counts: Optional[Mapping[str, float]] = None,
                  average_buffers: bool = False) -> nn.Module:
    """Average ``models`` **in place into the first one** and return it."""
    if len(models) == 0:
        raise ValueError("federated_avg needs at least one model")
    ids = list(models.keys())
    mods = [models[i] for i in ids]
    if len({id(m) for m in mods}) != len(mods):
        # The reference's local mode averages K aliases of one object and silently computes
        # theta * 2^(K-1) / K.  That is a bug, not a behaviour to keep.
        raise ValueError("federated_avg received aliased modules; every worker needs its own replica")
    counts = None if sample_counts is None else [sample_counts[i] for i in ids]
    w = normalized_weights(counts, len(mods))
    first = mods[0]
    plists = [list(m.parameters()) for m in mods]
    for j, p0 in enumerate(plists[0]):
        acc = p0.detach().to(torch.float32) * w[0]
        for k in range(1, len(mods)):
            acc += plists[k][j].detach().to(acc.device, torch.float32) * w[k]
        p0.copy_(acc.to(p0.dtype))
    if average_buffers:
        blists = [list(m.buffers()) for m in mods]
        for j, b0 in enumerate(blists[0]):
            if not b0.dtype.is_floating_point:
                continue
            acc = b0.detach().float() * w[0]
            for k in range(1, len(mods)):
                acc += blists[k][j].detach().to(acc.device).float() * w[k]
            b0.copy_(acc.to(b0.dtype))
    return first


@torch.no_grad()
def federated_avg_flat(flats: Sequence[torch.Tensor], sample_counts: Optional[Sequence[float]] = None
                       ) -> torch.Tensor:
    from .. import ops

    stacked = torch.stack([f.reshape(-1) for f in flats])
    w = normalized_weights(sample_counts, len(flats), device=stacked.device)
    return ops.fedavg_flat(stacked, w)
