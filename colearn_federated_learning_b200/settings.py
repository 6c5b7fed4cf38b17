"""Global device registry shared by the event-intake thread and the training thread.

Parity: reference ``settings.py:8-12`` keeps two bare module globals (``training_devices``,
``event_served``) mutated from the MQTT thread and the Timer thread without a lock
(SURVEY §2.8-9).  Here the same two names exist, but they are views onto a
:class:`DeviceRegistry` guarded by one re-entrant lock, so the snapshot taken when the
temporal window closes is atomic with respect to late TRAINING / NOT_READY events.
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import Any, Dict, Iterator, Optional


class DeviceRegistry:
    """Ordered ``id -> worker`` map plus the ``event_served`` counter, behind a lock.

    Insertion order is preserved because the encrypted trainer selects "the first two"
    registered devices (reference ``federated_coordinator.py:401-404``).
    """

    def __init__(self) -> None:
        self._lock = threading.RLock()
        self._devices: "OrderedDict[str, Any]" = OrderedDict()
        self._event_served = 0

    # -- counter -----------------------------------------------------------------
    @property
    def event_served(self) -> int:
        with self._lock:
            return self._event_served

    @event_served.setter
    def event_served(self, value: int) -> None:
        with self._lock:
            self._event_served = int(value)

    def serve_event(self) -> int:
        """Atomically ``event_served += 1``; returns the new value."""
        with self._lock:
            self._event_served += 1
            return self._event_served

    def reset_window(self) -> None:
        """Re-arm the temporal window (reference ``federated_coordinator.py:392,472,597``)."""
        self.event_served = 0

    # -- mapping protocol ----------------------------------------------------------
    def register(self, worker_id: str, worker: Any) -> None:
        with self._lock:
            self._devices[worker_id] = worker

    def remove(self, worker_id: str) -> Optional[Any]:
        with self._lock:
            return self._devices.pop(worker_id, None)

    def snapshot(self) -> "OrderedDict[str, Any]":
        """Atomic copy = the round membership (reference ``.copy()`` at fc.py:326,407,498)."""
        with self._lock:
            return OrderedDict(self._devices)

    def __contains__(self, worker_id: object) -> bool:
        with self._lock:
            return worker_id in self._devices

    def __getitem__(self, worker_id: str) -> Any:
        with self._lock:
            return self._devices[worker_id]

    def __setitem__(self, worker_id: str, worker: Any) -> None:
        self.register(worker_id, worker)

    def __delitem__(self, worker_id: str) -> None:
        with self._lock:
            del self._devices[worker_id]

    def __len__(self) -> int:
        with self._lock:
            return len(self._devices)

    def __iter__(self) -> Iterator[str]:
        return iter(self.snapshot())

    def get(self, worker_id: str, default: Any = None) -> Any:
        with self._lock:
            return self._devices.get(worker_id, default)

    def keys(self):
        return self.snapshot().keys()

    def values(self):
        return self.snapshot().values()

    def items(self):
        return self.snapshot().items()

    def copy(self) -> Dict[str, Any]:
        return self.snapshot()

    def clear(self) -> None:
        with self._lock:
            self._devices.clear()

    def __repr__(self) -> str:
        return f"DeviceRegistry({list(self.snapshot().keys())}, event_served={self.event_served})"


# Reference-compatible module-level surface -------------------------------------------------
registry = DeviceRegistry()
training_devices = registry  # dict-like


def init() -> DeviceRegistry:
    """Reset the global registry (reference ``settings.init()``, called at fc.py:102)."""
    global registry, training_devices
    registry = DeviceRegistry()
    training_devices = registry
    return registry


def __getattr__(name: str):  # PEP 562: ``settings.event_served`` reads through to the registry
    if name == "event_served":
        return registry.event_served
    raise AttributeError(name)


def set_event_served(value: int) -> None:
    registry.event_served = value
