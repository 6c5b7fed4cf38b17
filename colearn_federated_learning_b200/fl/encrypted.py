"""Encrypted (SMPC) training demo — reference ``starting_training_enc`` (fc.py:394-472),
``get_private_data_loaders`` (cf.py:257-277) and ``encrypted_training`` (cf.py:130-170).

Exactly two workers plus a crypto provider.  The first ``n_train_items`` samples of a shuffled
loader are fixed-precision-encoded and secret-shared, the model is shared, ``args.epochs``
passes of encrypted SGD run, then the model is reconstructed.  Unlike the reference (which
forgets to) the caller saves the decrypted model.
"""
from __future__ import annotations

import logging
import time
from typing import Any, Dict, List, Sequence, Tuple

import torch

from ..data import dataset_tensors
from ..models import MLPNet
from ..smpc import CryptoProvider, SharedMLP, fix_precision, float_precision, share

log = logging.getLogger(__name__)


def get_private_data_loaders(dataset, provider: CryptoProvider, n_train_items: int, batch_size: int = 1,
                             precision_fractional: int = 3, seed: int = 1) -> List[Tuple[Any, Any]]:
    """List of ``(shared data, shared target)`` batches (cf.py:269-277: the *first*
    ``n_train_items / batch_size`` batches of a shuffled loader)."""
    x, y = dataset_tensors(dataset)
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(len(x), generator=g)
    n_batches = int(n_train_items / batch_size)
    out = []
    for i in range(min(n_batches, (len(x) + batch_size - 1) // batch_size)):
        idx = order[i * batch_size:(i + 1) * batch_size]
        if len(idx) == 0:
            break
        out.append((share(fix_precision(x[idx], precision_fractional), provider),
                    share(fix_precision(y[idx].view(len(idx), -1), precision_fractional), provider)))
    return out


def encrypted_training(model: SharedMLP, loader, lr: float, epoch: int, log_interval: int = 30,
                       batch_size: int = 1) -> float:
    start = time.time()
    last = 0.0
    for batch_idx, (data, target) in enumerate(loader):
        loss = model.step(data, target, lr)
        if batch_idx % log_interval == 0:
            last = float(float_precision(loss.get()))
            print("Train Epoch: {} [{}/{} ({:.0f}%)]\tLoss: {:.6f}\tTime: {:.3f}s".format(
                epoch, batch_idx * batch_size, len(loader) * batch_size, 100.0 * batch_idx / max(1, len(loader)),
                last, time.time() - start))
    return last


def train_encrypted(model: MLPNet, dataset, worker_ids: Sequence[str], args) -> Dict[str, Any]:
    if len(worker_ids) != 2:
        raise ValueError("the encrypted demo runs on exactly two workers")
    if not isinstance(model, MLPNet):
        raise ValueError("encrypted training supports the MLP family only")
    provider = CryptoProvider(seed=args.seed)
    loader = get_private_data_loaders(dataset, provider, args.n_train_items_enc, args.batch_size,
                                      args.precision_fractional, args.seed)
    log.info("Encryption and distribution of the model...")
    shared = SharedMLP.from_module(model, provider)
    t0 = time.time()
    last = 0.0
    for epoch in range(args.epochs):
        last = encrypted_training(shared, loader, args.lr, epoch, args.log_interval, args.batch_size)
    shared.reveal_into(model)
    return {"workers": list(worker_ids), "batches": len(loader), "last_loss": last, "seconds": time.time() - t0,
            "triples": provider.triples_dealt, "comparisons": provider.comparisons}
