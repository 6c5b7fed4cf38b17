"""Dataset sharding across workers.

Contracts re-created from PySyft 0.2.x ([EXTERNAL]):

* ``dataset.federate(workers)`` — split **in order** into ``len(workers)`` contiguous shards of
  ``ceil(N / len(workers))`` samples; shard *i* lives on worker *i* (``fc.py:348-349``).
* ``FederatedDataLoader(fed, batch_size, shuffle=True)`` — iterate worker by worker; every batch
  comes from exactly one worker; shuffling is within a worker's shard (``fc.py:347-350``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence, Tuple

import torch

from .datasets import dataset_tensors


def shard_bounds(n: int, k: int) -> List[Tuple[int, int]]:
    """Contiguous ``ceil(n/k)``-sized shards; trailing shards may be short or empty."""
    size = math.ceil(n / k) if k > 0 else 0
    return [(min(i * size, n), min((i + 1) * size, n)) for i in range(k)]


@dataclass
class Shard:
    worker_id: str
    x: torch.Tensor
    y: torch.Tensor

    def __len__(self) -> int:
        return int(self.x.shape[0])


class FederatedDataset:
    def __init__(self, shards: "Dict[str, Shard]") -> None:
        self.shards = shards

    @property
    def workers(self) -> List[str]:
        return list(self.shards.keys())

    def __len__(self) -> int:
        return sum(len(s) for s in self.shards.values())

    def __getitem__(self, worker_id: str) -> Shard:
        return self.shards[worker_id]


def federate(dataset, worker_ids: Sequence[str], device: Optional[torch.device] = None) -> FederatedDataset:
    x, y = dataset_tensors(dataset, device)
    shards = {}
    for wid, (lo, hi) in zip(worker_ids, shard_bounds(len(x), len(worker_ids))):
        shards[wid] = Shard(wid, x[lo:hi], y[lo:hi])
    return FederatedDataset(shards)


class FederatedDataLoader:
    """Yields ``(worker_id, data, target)`` batches, worker by worker."""

    def __init__(self, fed: FederatedDataset, batch_size: int = 1, shuffle: bool = True,
                 seed: int = 1, drop_last: bool = False) -> None:
        self.fed, self.batch_size, self.shuffle, self.seed, self.drop_last = fed, batch_size, shuffle, seed, drop_last
        self._epoch = 0

    def __len__(self) -> int:
        f = math.floor if self.drop_last else math.ceil
        return sum(f(len(s) / self.batch_size) for s in self.fed.shards.values())

    def __iter__(self) -> Iterator[Tuple[str, torch.Tensor, torch.Tensor]]:
        g = torch.Generator().manual_seed(self.seed + self._epoch)
        self._epoch += 1
        for wid, shard in self.fed.shards.items():
            n = len(shard)
            order = torch.randperm(n, generator=g) if self.shuffle else torch.arange(n)
            order = order.to(shard.x.device)
            for lo in range(0, n, self.batch_size):
                idx = order[lo:lo + self.batch_size]
                if self.drop_last and len(idx) < self.batch_size:
                    break
                yield wid, shard.x[idx], shard.y[idx]
