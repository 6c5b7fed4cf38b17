"""Process-group bootstrap (one process per GPU; rank 0 = coordinator role).

``torchrun``/``torch.distributed.run`` provide RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR /
MASTER_PORT.  NCCL (or gloo on a CPU box) is used **only** for rendezvous, handle exchange and
barriers outside the timed region; the round path itself never calls a collective library.
"""
from __future__ import annotations

import datetime
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_distributed(backend: Optional[str] = None, timeout_s: int = 600) -> Tuple[int, int, torch.device]:
    """Initialise the default process group from the environment (no-op for WORLD_SIZE=1)."""
    rank, local_rank, world = env_world()
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if use_cuda else "gloo")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = device
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=timeout_s), **kwargs)
    return rank, world, device


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        try:
            dist.barrier()
        except Exception:  # noqa: BLE001
            pass
        dist.destroy_process_group()
