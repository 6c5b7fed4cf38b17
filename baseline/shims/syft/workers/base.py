"""Worker objects of the PySyft stand-in: an id, the registry of known workers, and the device its data lives on."""
from __future__ import annotations

from typing import Dict, Optional

import torch


class BaseWorker:
    def __init__(self, hook=None, id: str = "", **_kw) -> None:  # noqa: A002 - PySyft's keyword is `id`
        self.hook = hook
        self.id = id
        self.is_client_worker = True
        self._known_workers: Dict[str, "BaseWorker"] = {}
        self._objects: Dict[int, torch.Tensor] = {}
        self.device: Optional[torch.device] = None
        if hook is not None and getattr(hook, "local_worker", None) is not None:
            hook.local_worker._known_workers[id] = self

    def note_device(self, device: torch.device) -> None:
        if self.device is None or device.type == "cuda":
            self.device = device

    def search(self, *tags: str):
        return [t for t in self._objects.values() if set(tags) <= set(getattr(t, "tags", ()))]

    def close(self) -> None:
        pass

    def __repr__(self) -> str:
        return f"<{type(self).__name__} id:{self.id} #objects:{len(self._objects)}>"
