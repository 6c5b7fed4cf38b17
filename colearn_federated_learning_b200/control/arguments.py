"""Hyper-parameter object.

Parity: reference ``Arguments`` (``federated_coordinator.py:45-61``): batch_size 1,
test_batch_size 1024, epochs 1, federate_after_n_batches -1 (unlimited), lr 0.01, momentum 0.5
(unused by the reference's plain SGD), seed 1, log_interval 30, test_path = example CSV; setter
``set_federated_batches``.  Everything the reference hard-codes is a field here and is exposed
as a CLI flag by ``federated_coordinator.py`` with the reference value as default.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Any, Dict

DEFAULT_TEST_PATH = "./dataset_example/UNSW_2018_IoT_Botnet_Final_10_best_Training_1_1.csv"
ROUND_MODE_BATCHES = 1000  # reference: ``round > 1`` forces 1000 local batches (fc.py:533-535)


@dataclass
class Arguments:
    batch_size: int = 1
    test_batch_size: int = 1024
    epochs: int = 1
    federate_after_n_batches: int = -1
    lr: float = 0.01
    momentum: float = 0.5
    no_cuda: bool = False
    seed: int = 1
    log_interval: int = 30
    save_model: bool = False
    test_path: str = DEFAULT_TEST_PATH
    # -- extensions (not in the reference) -------------------------------------------------
    model: str = "ffnn"
    loss: str = "auto"             # auto | bce | sse | xent
    weighted: bool = False         # False = uniform FedAvg (reference), True = n_k / sum(n)
    server_lr: float = 1.0         # theta <- theta + server_lr * avg_delta (1.0 = plain FedAvg)
    backend: str = "auto"          # auto | fused | nccl | cpu
    dtype: str = "fp32"            # fp32 | bf16 (box mode: bf16 shadow of the broadcast for the tcgen05 consumers)
    save_every: int = 0            # >0: checkpoint every N rounds, not only after the last one
    synthetic: int = 0             # >0: generate this many synthetic UNSW-shaped rows
    dataset: str = "unsw"          # data.DATASET_REGISTRY entry that reads test_path (user datasets: data.register_dataset)
    n_train_items_enc: int = 1000  # fc.py:432
    precision_fractional: int = 3  # fc.py:433

    def set_federated_batches(self, batches: int) -> None:
        self.federate_after_n_batches = batches

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)
