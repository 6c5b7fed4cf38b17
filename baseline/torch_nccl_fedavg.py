"""Comparator: the same federated round written the ordinary way — stock PyTorch modules,
autograd, ``torch.optim.SGD`` (cuBLAS/cuDNN underneath) and NCCL collectives.

It exists because the reference has no GPU/NCCL build of its own and cannot be installed
offline: this file is "the baseline, not the product" (SURVEY §6.3).
It deliberately imports **nothing** from ``colearn_federated_learning_b200.ops`` /
``.parallel`` — only the ``nn.Module`` definitions and the synthetic-data generator — so that
none of this repo's kernels or engine sit on its path.

Round semantics (identical to the fused engine): ``dist.broadcast(θ)`` → every selected rank
runs ``epochs`` of shuffled mini-batch SGD on its shard → ``dist.reduce(w_k·θ_k)`` → θ ← Σ.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import parameters_to_vector, vector_to_parameters


def _loss(out: torch.Tensor, y: torch.Tensor, loss: str) -> torch.Tensor:
    if loss == "xent":
        return F.cross_entropy(out.float(), y.view(-1).long())
    if loss == "bce":
        return F.binary_cross_entropy(out, y.view_as(out))
    return ((out - y.view_as(out)) ** 2).sum()


class TorchNcclFedAvg:
    def __init__(self, model: nn.Module, device: torch.device, loss: str = "xent", batch_size: int = 1,
                 lr: float = 0.01, epochs: int = 1, max_batches: int = -1, weighted: bool = True,
                 bf16_autocast: bool = False, seed: int = 1) -> None:
        self.model = model.to(device)
        self.device = device
        self.loss, self.batch_size, self.lr, self.epochs, self.max_batches = loss, batch_size, lr, epochs, max_batches
        self.weighted = weighted
        self.bf16 = bf16_autocast
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.gen = torch.Generator(device="cpu").manual_seed(seed + self.rank)
        self.theta = parameters_to_vector(self.model.parameters()).detach().clone()
        self.counts: List[int] = [0] * self.world

    def set_local_data(self, x: torch.Tensor, y: torch.Tensor) -> None:
        self.x, self.y = x.to(self.device), y.to(self.device)
        n = torch.tensor([self.x.shape[0]], device=self.device, dtype=torch.float32)
        if self.world > 1:
            allc = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(allc, n)
            self.counts = [int(c.item()) for c in allc]
        else:
            self.counts = [int(n.item())]

    def _local_fit(self) -> torch.Tensor:
        opt = torch.optim.SGD(self.model.parameters(), lr=self.lr)
        n = self.x.shape[0]
        it = 0
        last = torch.zeros((), device=self.device)
        for _ in range(self.epochs):
            order = torch.randperm(n, generator=self.gen).to(self.device)
            for lo in range(0, n, self.batch_size):
                idx = order[lo:lo + self.batch_size]
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
                    out = self.model(self.x[idx])
                last = _loss(out, self.y[idx], self.loss)
                last.backward()
                opt.step()
                it += 1
                if self.max_batches > 0 and it >= self.max_batches:
                    return last.detach()
        return last.detach()

    def run_round(self, mask: Optional[int] = None, host_inputs: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> float:
        full = (1 << self.world) - 1
        mask = full if mask is None else mask & full
        if host_inputs is not None:
            self.x.copy_(host_inputs[0], non_blocking=True)
            self.y.copy_(host_inputs[1].view_as(self.y), non_blocking=True)
        if self.world > 1:
            dist.broadcast(self.theta, src=0)                                   # broadcast leg
        selected = (mask >> self.rank) & 1
        sel = [k for k in range(self.world) if (mask >> k) & 1]
        tot = float(sum(self.counts[k] for k in sel)) if self.weighted else float(len(sel))
        w = ((self.counts[self.rank] if self.weighted else 1.0) / tot) if selected else 0.0
        loss = torch.zeros((), device=self.device)
        if selected:
            with torch.no_grad():
                vector_to_parameters(self.theta, self.model.parameters())
            loss = self._local_fit()
            contrib = parameters_to_vector(self.model.parameters()).detach() * w
        else:
            contrib = torch.zeros_like(self.theta)
        if self.world > 1:
            dist.reduce(contrib, dst=0)                                          # gather + FedAvg reduce
        if self.rank == 0:
            self.theta.copy_(contrib)                                            # server apply (lr_s = 1)
        return float(loss)                                                       # D2H read of the result
