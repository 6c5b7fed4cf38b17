"""Publish/subscribe signalling plane (replaces the MQTT broker + paho client of the reference).

The reference's control plane is MQTT 3.1 via paho, qos 0, a single topic (SURVEY §5;
``federated_coordinator.py:92,294-300``, ``remote_worker.py:65-68,110``).  On a single NVSwitch
box there is no broker to talk to, so the bus is:

* :class:`InProcessBroker` — topic → subscribers in one process.  This *is* the multi-node fake
  used by the tests (it mirrors the reference's VirtualWorker idea) and carries the fault-
  injection hooks (drop / delay / duplicate / rewrite) required by SURVEY §5.
* :class:`TcpBroker` / ``BusClient(transport="tcp")`` — the same protocol over a localhost TCP
  socket (newline-delimited JSON frames) so that ``remote_worker.py`` processes and
  ``federated_coordinator.py`` can find each other exactly like they do through mosquitto.
* :class:`BusClient` — a paho-shaped client (``connect / subscribe / publish / loop_forever /
  loop_start / loop_stop / disconnect`` and the ``on_connect / on_message / on_publish``
  callbacks with the paho signatures) so the Coordinator reads like the reference's.

MQTT topic filters ``+`` (one level) and ``#`` (rest) are honoured.
"""
from __future__ import annotations

import base64
import json
import logging
import queue
import socket
import socketserver
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple


log = logging.getLogger("colearn.bus")


@dataclass
class Message:
    """Shape-compatible with ``paho.mqtt.client.MQTTMessage`` (topic / payload / qos / mid)."""

    topic: str
    payload: bytes
    qos: int = 0
    retain: bool = False
    mid: int = 0
    timestamp: float = field(default_factory=time.time)


def topic_matches(pattern: str, topic: str) -> bool:
    """MQTT topic-filter matching (``+`` single level, ``#`` multi level)."""
    p_parts, t_parts = pattern.split("/"), topic.split("/")
    for i, p in enumerate(p_parts):
        if p == "#":
            return True
        if i >= len(t_parts):
            return False
        if p != "+" and p != t_parts[i]:
            return False
    return len(p_parts) == len(t_parts)


# A fault hook sees every message before delivery and returns the list of (delay_s, Message)
# deliveries to perform: [] drops it, two entries duplicate it, delay > 0 reorders it.
FaultHook = Callable[[Message], List[Tuple[float, Message]]]


class InProcessBroker:
    """Thread-safe topic router with fault injection."""

    def __init__(self) -> None:
        self._lock = threading.RLock()
        self._subs: List[Tuple[str, "BusClient"]] = []
        self._hooks: List[FaultHook] = []
        self._mid = 0
        self.delivered = 0
        self.dropped = 0
        self._timers: List[threading.Timer] = []

    # -- subscription management -----------------------------------------------------
    def add_subscription(self, pattern: str, client: "BusClient") -> None:
        with self._lock:
            self._subs.append((pattern, client))

    def remove_client(self, client: "BusClient") -> None:
        with self._lock:
            self._subs = [(p, c) for (p, c) in self._subs if c is not client]

    # -- fault injection ---------------------------------------------------------------
    def add_fault_hook(self, hook: FaultHook) -> None:
        with self._lock:
            self._hooks.append(hook)

    def clear_fault_hooks(self) -> None:
        with self._lock:
            self._hooks.clear()

    def inject_drop(self, predicate: Callable[[Message], bool]) -> None:
        self.add_fault_hook(lambda m: [] if predicate(m) else [(0.0, m)])

    def inject_delay(self, predicate: Callable[[Message], bool], seconds: float) -> None:
        self.add_fault_hook(lambda m: [(seconds, m)] if predicate(m) else [(0.0, m)])

    def inject_duplicate(self, predicate: Callable[[Message], bool]) -> None:
        self.add_fault_hook(lambda m: [(0.0, m), (0.0, m)] if predicate(m) else [(0.0, m)])

    def inject_from_spec(self, spec: str) -> None:
        """Install a fault hook from a CLI string (``federated_coordinator.py --inject``):
        ``drop:<regex>`` | ``dup:<regex>`` | ``delay:<seconds>:<regex>`` — the regex is searched in the utf-8 payload."""
        import re
        kind, _, rest = spec.partition(":")
        if kind == "delay":
            secs, _, pattern = rest.partition(":")
            rx = re.compile(pattern or ".")
            self.inject_delay(lambda m: bool(rx.search(m.payload.decode("utf-8", "replace"))), float(secs))
        elif kind in ("drop", "dup"):
            rx = re.compile(rest or ".")
            pred = lambda m: bool(rx.search(m.payload.decode("utf-8", "replace")))  # noqa: E731
            (self.inject_drop if kind == "drop" else self.inject_duplicate)(pred)
        else:
            raise ValueError(f"unknown fault spec {spec!r} (use drop:<re>, dup:<re> or delay:<s>:<re>)")

    # -- publish -------------------------------------------------------------------------
    def publish(self, topic: str, payload, qos: int = 0) -> int:
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        with self._lock:
            self._mid += 1
            mid = self._mid
            hooks = list(self._hooks)
        deliveries: List[Tuple[float, Message]] = [(0.0, Message(topic, bytes(payload), qos, mid=mid))]
        for hook in hooks:
            nxt: List[Tuple[float, Message]] = []
            for delay, msg in deliveries:
                for d2, m2 in hook(msg):
                    nxt.append((delay + d2, m2))
            deliveries = nxt
        if not deliveries:
            with self._lock:
                self.dropped += 1
        for delay, msg in deliveries:
            if delay > 0:
                t = threading.Timer(delay, self._deliver, args=(msg,))
                t.daemon = True
                with self._lock:
                    self._timers = [x for x in self._timers if x.is_alive()]   # fired timers are not kept around
                    self._timers.append(t)
                t.start()
            else:
                self._deliver(msg)
        return mid

    def _deliver(self, msg: Message) -> None:
        with self._lock:
            targets = [c for (p, c) in self._subs if topic_matches(p, msg.topic)]
            self.delivered += len(targets)
        for c in targets:
            c._enqueue(msg)

    def shutdown(self) -> None:
        with self._lock:
            for t in self._timers:
                t.cancel()
            self._timers.clear()


_default_broker: Optional[InProcessBroker] = None
_default_lock = threading.Lock()


def default_broker() -> InProcessBroker:
    global _default_broker
    with _default_lock:
        if _default_broker is None:
            _default_broker = InProcessBroker()
        return _default_broker


def reset_default_broker() -> InProcessBroker:
    global _default_broker
    with _default_lock:
        if _default_broker is not None:
            _default_broker.shutdown()
        _default_broker = InProcessBroker()
        return _default_broker


# ---------------------------------------------------------------------------------------------
# TCP transport (cross-process, localhost) — newline-delimited JSON frames:
#   {"op":"sub","topic":...} / {"op":"pub","topic":...,"payload":<b64>,"qos":0}
#   broker -> client: {"op":"msg","topic":...,"payload":<b64>,"qos":0,"mid":n}
# ---------------------------------------------------------------------------------------------
class _TcpBridgeClient:
    """Broker-side proxy for one TCP connection; quacks like a BusClient for ``_enqueue``."""

    def __init__(self, wfile, lock: threading.Lock) -> None:
        self._wfile = wfile
        self._lock = lock
        self.alive = True

    def _enqueue(self, msg: Message) -> None:
        if not self.alive:
            return
        frame = json.dumps({"op": "msg", "topic": msg.topic, "qos": msg.qos, "mid": msg.mid,
                            "payload": base64.b64encode(msg.payload).decode("ascii")}) + "\n"
        try:
            with self._lock:
                self._wfile.write(frame.encode("utf-8"))
                self._wfile.flush()
        except OSError:
            self.alive = False


class _TcpHandler(socketserver.StreamRequestHandler):
    def handle(self) -> None:  # one thread per connection
        broker: InProcessBroker = self.server.broker  # type: ignore[attr-defined]
        proxy = _TcpBridgeClient(self.wfile, threading.Lock())
        try:
            for raw in self.rfile:
                try:
                    frame = json.loads(raw.decode("utf-8"))
                except ValueError:
                    continue
                if not isinstance(frame, dict) or not isinstance(frame.get("topic", ""), str):
                    continue
                op = frame.get("op")
                try:
                    if op == "sub" and "topic" in frame:
                        broker.add_subscription(frame["topic"], proxy)  # type: ignore[arg-type]
                    elif op == "pub" and "topic" in frame:
                        broker.publish(frame["topic"], base64.b64decode(frame.get("payload", "")),
                                       int(frame.get("qos", 0)))
                    elif op == "ping":
                        proxy._enqueue(Message("$SYS/pong", b""))
                except (ValueError, TypeError):        # bad base64 / qos: drop the frame, keep the connection
                    continue
        finally:
            proxy.alive = False
            broker.remove_client(proxy)  # type: ignore[arg-type]


class _ThreadedTCPServer(socketserver.ThreadingMixIn, socketserver.TCPServer):
    allow_reuse_address = True
    daemon_threads = True


class TcpBroker:
    """A tiny stand-in for mosquitto on 127.0.0.1 (default port 1883 like MQTT)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 1883,
                 broker: Optional[InProcessBroker] = None) -> None:
        self.broker = broker or InProcessBroker()
        self._server = _ThreadedTCPServer((host, port), _TcpHandler)
        self._server.broker = self.broker  # type: ignore[attr-defined]
        self.host, self.port = self._server.server_address[:2]
        self._thread: Optional[threading.Thread] = None

    def start(self) -> "TcpBroker":
        self._thread = threading.Thread(target=self._server.serve_forever, name="bus-broker", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._server.shutdown()
        self._server.server_close()
        self.broker.shutdown()

    def __enter__(self) -> "TcpBroker":
        return self.start()

    def __exit__(self, *exc) -> None:
        self.stop()


# ---------------------------------------------------------------------------------------------
class BusClient:
    """paho-shaped pub/sub client.  Subclass it and override ``on_message`` etc., or assign
    callables to those attributes — both styles work, as with paho."""

    def __init__(self, client_id: str = "", broker: Optional[InProcessBroker] = None,
                 transport: str = "inproc") -> None:
        self.client_id = client_id
        self._transport = transport
        self._broker = broker
        self._queue: "queue.Queue[Optional[Message]]" = queue.Queue()
        self._running = threading.Event()
        self._loop_thread: Optional[threading.Thread] = None
        self._sock: Optional[socket.socket] = None
        self._sock_lock = threading.Lock()
        self._reader: Optional[threading.Thread] = None
        self._connected = False
        self._mid = 0
        # loop_start()/loop_forever() survive a raising callback (logged); drain()/loop() called directly from
        # tests re-raise unless this is set
        self.suppress_callback_errors = False

    # Default callbacks (paho signatures) — overridable --------------------------------
    def on_connect(self, client, userdata, flags, rc) -> None:  # noqa: D401
        pass

    def on_message(self, client, userdata, msg: Message) -> None:
        pass

    def on_publish(self, client, userdata, mid) -> None:
        pass

    # Connection ---------------------------------------------------------------------------
    def connect(self, host: str = "localhost", port: int = 1883, keepalive: int = 60) -> int:
        if self._transport == "tcp":
            addr = "127.0.0.1" if host in ("localhost", "") else host
            self._sock = socket.create_connection((addr, port), timeout=10)
            self._sock.settimeout(None)
            self._reader = threading.Thread(target=self._tcp_reader, name="bus-reader", daemon=True)
            self._reader.start()
        else:
            if self._broker is None:
                self._broker = default_broker()
        self._connected = True
        self.on_connect(self, None, {}, 0)
        return 0

    def _tcp_reader(self) -> None:
        assert self._sock is not None
        f = self._sock.makefile("rb")
        try:
            for raw in f:
                try:
                    frame = json.loads(raw.decode("utf-8"))
                except ValueError:
                    continue
                if frame.get("op") == "msg":
                    self._enqueue(Message(frame["topic"], base64.b64decode(frame.get("payload", "")),
                                          int(frame.get("qos", 0)), mid=int(frame.get("mid", 0))))
        except OSError:
            pass

    def _send(self, frame: dict) -> None:
        assert self._sock is not None
        data = (json.dumps(frame) + "\n").encode("utf-8")
        with self._sock_lock:
            self._sock.sendall(data)

    def subscribe(self, topic: str, qos: int = 0):
        if self._transport == "tcp":
            self._send({"op": "sub", "topic": topic, "qos": qos})
        else:
            assert self._broker is not None, "connect() first"
            self._broker.add_subscription(topic, self)
        return (0, 0)

    def publish(self, topic: str, payload=b"", qos: int = 0, retain: bool = False):
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        if self._transport == "tcp":
            self._send({"op": "pub", "topic": topic, "qos": qos,
                        "payload": base64.b64encode(payload).decode("ascii")})
            self._mid += 1
            mid = self._mid
        else:
            assert self._broker is not None, "connect() first"
            mid = self._broker.publish(topic, payload, qos)
        self.on_publish(self, None, mid)
        return mid

    def _enqueue(self, msg: Message) -> None:
        self._queue.put(msg)

    # Network loop ---------------------------------------------------------------------------
    def loop(self, timeout: float = 0.1) -> int:
        """Process at most one pending message; returns the number handled."""
        try:
            msg = self._queue.get(timeout=timeout)
        except queue.Empty:
            return 0
        if msg is None:
            return 0
        self._dispatch(msg)
        return 1

    def _dispatch(self, msg: Message) -> None:
        """A callback that raises must not take the network loop (and with it the whole control plane) down:
        log it and keep serving, like a broker connection would."""
        try:
            self.on_message(self, None, msg)
        except Exception:  # noqa: BLE001
            if self.suppress_callback_errors:
                log.exception("on_message raised for topic %r (client %r); message dropped", msg.topic, self.client_id)
            else:
                raise

    def drain(self) -> int:
        """Synchronously deliver everything currently queued (handy in tests)."""
        n = 0
        while True:
            try:
                msg = self._queue.get_nowait()
            except queue.Empty:
                return n
            if msg is not None:
                self._dispatch(msg)
                n += 1

    def loop_forever(self) -> None:
        self._running.set()
        prev, self.suppress_callback_errors = self.suppress_callback_errors, True
        try:
            while self._running.is_set():
                self.loop(timeout=0.1)
        finally:
            self.suppress_callback_errors = prev

    def loop_start(self) -> None:
        if self._loop_thread is None or not self._loop_thread.is_alive():
            self._loop_thread = threading.Thread(target=self.loop_forever, name="bus-loop", daemon=True)
            self._loop_thread.start()
            # make sure the loop is live before returning so an immediate publish is seen
            while not self._running.is_set():
                time.sleep(0.001)

    def loop_stop(self) -> None:
        self._running.clear()
        self._queue.put(None)
        if self._loop_thread is not None and self._loop_thread is not threading.current_thread():
            self._loop_thread.join(timeout=5)
        self._loop_thread = None

    def disconnect(self) -> None:
        self.loop_stop()
        if self._transport == "tcp" and self._sock is not None:
            try:
                self._sock.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
            self._sock.close()
            self._sock = None
        elif self._broker is not None:
            self._broker.remove_client(self)
        self._connected = False


def make_client(client_id: str = "", host: Optional[str] = None, port: int = 1883,
                broker: Optional[InProcessBroker] = None) -> BusClient:
    """``host=None`` → in-process bus; otherwise a TCP client for :class:`TcpBroker`."""
    if host is None:
        return BusClient(client_id, broker=broker, transport="inproc")
    return BusClient(client_id, transport="tcp")
