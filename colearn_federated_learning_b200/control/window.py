"""Temporal-window scheduler.

Behaviour (reference ``federated_coordinator.py:180-225`` and paper §3.2): the *first* TRAINING
event since the last reset arms a one-shot timer of ``window`` seconds; every device that
announces TRAINING before it fires is collected; when it fires the training function runs on an
atomic **snapshot** of the registry.  Devices that arrive while training is running stay
registered and ride the next window ("keep" policy).  NOT_READY inside the window removes the
device.  The window re-arms when the training function resets ``event_served`` to 0
(fc.py:392,472,597) — which here happens in a ``finally`` so a crashing trainer can never wedge
the scheduler (the reference only guarantees this for remote mode, fc.py:595-597).

Unlike the reference, time is injectable: production uses ``threading.Timer`` (same as the
reference), tests use :class:`FakeClock` and advance it deterministically.
"""
from __future__ import annotations

import heapq
import itertools
import logging
import threading
from typing import Any, Callable, Dict, List, Optional, Tuple

from ..settings import DeviceRegistry

log = logging.getLogger(__name__)


class TimerHandle:
    def cancel(self) -> None:  # pragma: no cover - interface
        raise NotImplementedError


class ThreadingTimerFactory:
    """Real time: one ``threading.Timer`` per armed window (reference fc.py:202-203,224-225)."""

    def start(self, delay: float, fn: Callable[[], None]) -> threading.Timer:
        t = threading.Timer(delay, fn)
        t.daemon = True
        t.start()
        return t


class _FakeHandle(TimerHandle):
    def __init__(self) -> None:
        self.cancelled = False

    def cancel(self) -> None:
        self.cancelled = True


class FakeClock:
    """Deterministic timer source: callbacks run synchronously inside :meth:`advance`."""

    def __init__(self) -> None:
        self.now = 0.0
        self._heap: List[Tuple[float, int, _FakeHandle, Callable[[], None]]] = []
        self._seq = itertools.count()

    def start(self, delay: float, fn: Callable[[], None]) -> _FakeHandle:
        h = _FakeHandle()
        heapq.heappush(self._heap, (self.now + delay, next(self._seq), h, fn))
        return h

    def advance(self, seconds: float) -> int:
        """Move time forward, firing due timers in order; returns how many fired."""
        target = self.now + seconds
        fired = 0
        while self._heap and self._heap[0][0] <= target:
            when, _, h, fn = heapq.heappop(self._heap)
            self.now = when
            if not h.cancelled:
                fn()
                fired += 1
        self.now = target
        return fired

    @property
    def pending(self) -> int:
        return sum(1 for (_, _, h, _) in self._heap if not h.cancelled)


TrainFn = Callable[[Dict[str, Any]], Any]


class TemporalWindow:
    """State machine: IDLE --first TRAINING--> COLLECTING --timer--> TRAINING --done--> IDLE."""

    IDLE, COLLECTING, TRAINING = "IDLE", "COLLECTING", "TRAINING"

    def __init__(self, registry: DeviceRegistry, window: float, train_fn: TrainFn,
                 timer_factory=None, lower_bound: int = 1,
                 rearm_if_pending: bool = False) -> None:
        self.registry = registry
        self.window = float(window)
        self.train_fn = train_fn
        self.timers = timer_factory or ThreadingTimerFactory()
        self.lower_bound = lower_bound
        # Reference behaviour: devices that joined during training wait for the *next TRAINING
        # event* to arm a window.  ``rearm_if_pending=True`` arms immediately instead.
        self.rearm_if_pending = rearm_if_pending
        self._lock = threading.RLock()
        self._state = self.IDLE
        self._handle = None
        self.windows_fired = 0
        self.last_result: Any = None
        self.last_error: Optional[BaseException] = None
        self.history: List[Dict[str, Any]] = []

    @property
    def state(self) -> str:
        with self._lock:
            return self._state

    # -- events -----------------------------------------------------------------------------
    def on_training(self, worker_id: str, worker: Any) -> bool:
        """Register the device; arm the timer iff this is the first event since the last reset.
        Returns True when a new window was armed."""
        with self._lock:
            n = self.registry.serve_event()           # fc.py:181
            self.registry.register(worker_id, worker)  # fc.py:182
            if n == 1:                                  # fc.py:187
                self._arm()
                return True
            return False

    def on_not_ready(self, worker_id: str) -> Optional[Any]:
        """Remove a device that withdrew before training started (fc.py:267-281)."""
        return self.registry.remove(worker_id)

    def _arm(self) -> None:
        self._state = self.COLLECTING
        log.info("Timer starting")
        self._handle = self.timers.start(self.window, self._fire)

    def cancel(self) -> bool:
        """Cancel a pending window (Ctrl-C path, fc.py:301-305).  Only effective while waiting."""
        with self._lock:
            if self._state == self.COLLECTING and self._handle is not None:
                self._handle.cancel()
                self._state = self.IDLE
                self.registry.reset_window()
                return True
            return False

    # -- timer body -----------------------------------------------------------------------------
    def _fire(self) -> None:
        with self._lock:
            if self._state != self.COLLECTING:
                return
            self._state = self.TRAINING
            snapshot = self.registry.snapshot()  # round membership
            self.windows_fired += 1
        record: Dict[str, Any] = {"window": self.windows_fired, "members": list(snapshot.keys())}
        try:
            if len(snapshot) >= self.lower_bound:
                self.last_result = self.train_fn(snapshot)
                record["trained"] = True
            else:
                log.info("No behaviour defined for the number of devices achieved")
                self.last_result = None
                record["trained"] = False
            self.last_error = None
        except BaseException as e:  # noqa: BLE001 - must never wedge the scheduler
            self.last_error = e
            record["error"] = repr(e)
            log.exception("training function failed")
        finally:
            with self._lock:
                self.registry.reset_window()  # "Restarting window"
                self._state = self.IDLE
                self.history.append(record)
                if self.rearm_if_pending and len(self.registry) > 0:
                    self.registry.serve_event()
                    self._arm()
