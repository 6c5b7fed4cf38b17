"""Checkpoint / resume.

Parity: the reference saves ``torch.save(model.state_dict(), './test.pth')`` once after the last
round (fc.py:381,585) and loads it — if present — at the start of every training window and per
inference request (fc.py:235-240,331-337,414-423,505-517).  Same contract here (same keys,
loadable by stock torch) plus: atomic rename, an optional JSON sidecar (round index, sample
counters, seed) and ``save_every``.  The encrypted trainer saves too (the reference forgets to,
SURVEY §2.8-6).
"""
from __future__ import annotations

import json
import logging
import os
import tempfile
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

log = logging.getLogger(__name__)

DEFAULT_PATH = "./test.pth"  # fc.py:123


def save_state_dict(state: Dict[str, torch.Tensor], path: str = DEFAULT_PATH,
                    meta: Optional[Dict[str, Any]] = None) -> str:
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    fd, tmp = tempfile.mkstemp(prefix="ckpt-tmp-", suffix=".pth", dir=d)
    os.close(fd)
    try:
        torch.save({k: v.detach().cpu() for k, v in state.items()}, tmp)
        os.replace(tmp, path)  # atomic on POSIX
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    if meta is not None:
        with open(path + ".json.tmp", "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        os.replace(path + ".json.tmp", path + ".json")
    return path


def save_model(model: nn.Module, path: str = DEFAULT_PATH, meta: Optional[Dict[str, Any]] = None) -> str:
    return save_state_dict(model.state_dict(), path, meta)


def load_or_init(model: nn.Module, path: str = DEFAULT_PATH, strict: bool = True) -> bool:
    """Load-if-exists else keep the random init.  Returns True when a checkpoint was loaded."""
    if not os.path.exists(path):
        log.info("No existing model")
        return False
    log.info("Found a model..")
    state = torch.load(path, map_location="cpu", weights_only=True)
    model.load_state_dict(state, strict=strict)
    return True


def load_meta(path: str = DEFAULT_PATH) -> Dict[str, Any]:
    try:
        with open(path + ".json") as f:
            return json.load(f)
    except (FileNotFoundError, ValueError):
        return {}


def checkpoint_compatible(model: nn.Module, path: str) -> bool:
    """True iff ``path`` holds exactly this architecture (guards the reference's INFERENCE bug of
    loading an FFNN checkpoint into TestingRemote)."""
    if not os.path.exists(path):
        return False
    state = torch.load(path, map_location="cpu", weights_only=True)
    own = model.state_dict()
    return set(state) == set(own) and all(state[k].shape == own[k].shape for k in own)
