"""Plain-PyTorch stand-in for the parts of PySyft 0.2.x the reference imports (see ../README.md).

Not PySyft, not part of the product, imports nothing from ``colearn_federated_learning_b200``."""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import torch
from torch import nn
from torch.utils.data import DataLoader, Dataset

from .workers.virtual import VirtualWorker  # noqa: F401  (sy.VirtualWorker)
from .workers.base import BaseWorker

__version__ = "0.2.x-standin"


class TorchHook:
    """``sy.TorchHook(torch)``: creates the local worker and teaches tensors / modules / datasets the handful of methods
    the reference calls on them (``send``, ``get``, ``location``, ``federate``, ``tag``)."""

    def __init__(self, torch_module=torch, local_worker: "BaseWorker | None" = None, is_client: bool = True) -> None:
        self.torch = torch_module
        self.local_worker = local_worker or VirtualWorker(None, "me")
        self.local_worker.hook = self
        self.local_worker._known_workers["me"] = self.local_worker
        self.local_worker.is_client_worker = is_client
        _install_methods()


def _tensor_send(self: torch.Tensor, *locations: BaseWorker, **_kw) -> torch.Tensor:
    loc = locations[0]
    out = self.detach() if not self.requires_grad else self
    out.location = loc
    loc.note_device(out.device)
    loc._objects[id(out)] = out
    return out


def _tensor_get(self: torch.Tensor, *_a, **_kw) -> torch.Tensor:
    loc = getattr(self, "location", None)
    if loc is not None:
        loc._objects.pop(id(self), None)
        self.location = None
    return self


def _tensor_tag(self: torch.Tensor, *tags: str) -> torch.Tensor:
    self.tags = set(getattr(self, "tags", ())) | set(tags)
    return self


def _module_send(self: nn.Module, *locations: BaseWorker, **_kw) -> nn.Module:
    loc = locations[0]
    dev = loc.device
    if dev is not None:
        p = next(self.parameters(), None)
        if p is not None and p.device != dev:
            self.to(dev)             # in place: parameter identities (and the optimizer's references) survive
    self.location = loc
    return self


def _module_get(self: nn.Module, *_a, **_kw) -> nn.Module:
    self.location = None
    return self


class BaseDataset(Dataset):
    """``sy.BaseDataset(data, targets)``: a pair of stacked tensors that can be sent to a worker."""

    def __init__(self, data: torch.Tensor, targets: torch.Tensor, transform=None) -> None:
        self.data, self.targets, self.transform_ = data, targets, transform

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, index):
        x = self.data[index]
        if self.transform_ is not None:
            x = self.transform_(x)
        return x, self.targets[index]

    def send(self, worker: BaseWorker) -> "BaseDataset":
        self.data = _tensor_send(self.data, worker)
        self.targets = _tensor_send(self.targets, worker)
        return self

    @property
    def location(self):
        return getattr(self.data, "location", None)


class FederatedDataset:
    """What ``dataset.federate(workers)`` returns: one ``BaseDataset`` per worker."""

    def __init__(self, datasets: Sequence[BaseDataset]) -> None:
        self.datasets: Dict[str, BaseDataset] = {}
        for ds in datasets:
            self.datasets[ds.location.id] = ds

    @property
    def workers(self) -> List[str]:
        return list(self.datasets.keys())

    def __getitem__(self, worker_id: str) -> BaseDataset:
        return self.datasets[worker_id]

    def __len__(self) -> int:
        return sum(len(d) for d in self.datasets.values())


def _dataset_federate(self: Dataset, workers: Sequence[BaseWorker]) -> FederatedDataset:
    """Contiguous shards of ``ceil(N / K)`` samples, shard *i* on worker *i* (every sample goes through the dataset's
    ``__getitem__`` and transform once, as a ``DataLoader`` pass does)."""
    n = len(self)
    per = int(math.ceil(n / max(1, len(workers))))
    loader = DataLoader(self, batch_size=per)
    shards = []
    for i, (data, targets) in enumerate(loader):
        w = workers[i % len(workers)]
        shards.append(BaseDataset(data, targets).send(w))
    return FederatedDataset(shards)


class FederatedDataLoader:
    """``sy.FederatedDataLoader(federated_dataset, batch_size, shuffle)``: iterates worker by worker; every batch comes
    from exactly one worker and carries it as ``.location``; ``shuffle`` permutes within a worker's shard."""

    def __init__(self, federated_dataset: FederatedDataset, batch_size: int = 8, shuffle: bool = False,
                 drop_last: bool = False, **_kw) -> None:
        self.federated_dataset = federated_dataset
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.workers = federated_dataset.workers

    def __len__(self) -> int:
        total = 0
        for w in self.workers:
            n = len(self.federated_dataset[w])
            total += n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size
        return total

    def __iter__(self):
        bs = self.batch_size
        for w in self.workers:
            ds = self.federated_dataset[w]
            n = len(ds)
            order = torch.randperm(n) if self.shuffle else torch.arange(n)
            stop = (n // bs) * bs if self.drop_last else n
            for lo in range(0, stop, bs):
                idx = order[lo:lo + bs].to(ds.data.device)
                data, target = ds.data[idx], ds.targets[idx]
                data.location = target.location = ds.location
                yield data, target


class TrainConfig:
    """Placeholder: the remote (websocket) path of the reference is not covered by this stand-in."""

    def __init__(self, *a, **kw) -> None:
        raise NotImplementedError("stand-in syft: TrainConfig / remote fit is not implemented (local VirtualWorker path only)")


_INSTALLED = False


def _install_methods() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    torch.Tensor.send = _tensor_send
    torch.Tensor.get = _tensor_get
    torch.Tensor.tag = _tensor_tag
    nn.Module.send = _module_send
    nn.Module.get = _module_get
    Dataset.federate = _dataset_federate
    _INSTALLED = True
