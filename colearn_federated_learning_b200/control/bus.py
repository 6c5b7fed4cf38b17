"""Publish/subscribe signalling plane (replaces the MQTT broker + paho client of the reference).

The reference's control plane is MQTT 3.1 via paho, qos 0, a single topic (SURVEY §5;
``federated_coordinator.py:92,294-300``, ``remote_worker.py:65-68,110``).  On a single NVSwitch
box there is no broker to talk to, so the bus is:

* :class:`InProcessBroker` — topic → subscribers in one process.  This *is* the multi-node fake
  used by the tests (it mirrors the reference's VirtualWorker idea) and carries the fault-
  injection hooks (drop / delay / duplicate / rewrite) required by SURVEY §5.
* :class:`TcpBroker` / ``BusClient(transport="tcp")`` — the same bus over TCP speaking **MQTT 3.1.1**
  (``control/mqtt.py``): ``remote_worker.py`` processes and ``federated_coordinator.py`` find each other exactly
  like they do through mosquitto, ``mosquitto_pub`` / paho clients can publish to the embedded broker, and the
  clients can equally connect to a real mosquitto.  Retained messages, last-will and keep-alive are honoured.
* :class:`BusClient` — a paho-shaped client (``connect / subscribe / publish / loop_forever /
  loop_start / loop_stop / disconnect`` and the ``on_connect / on_message / on_publish``
  callbacks with the paho signatures) so the Coordinator reads like the reference's.

MQTT topic filters ``+`` (one level) and ``#`` (rest) are honoured.
"""
from __future__ import annotations

import logging
import queue
import socket
import socketserver
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from . import mqtt


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
        self._retained: dict = {}     # topic -> Message (MQTT retained messages)

    # -- subscription management -----------------------------------------------------
    def add_subscription(self, pattern: str, client: "BusClient") -> None:
        with self._lock:
            if (pattern, client) not in self._subs:
                self._subs.append((pattern, client))
            retained = [m for t, m in self._retained.items() if topic_matches(pattern, t)]
        for m in retained:                      # a new subscriber gets the last retained message per topic
            client._enqueue(m)

    def remove_subscription(self, pattern: str, client: "BusClient") -> None:
        with self._lock:
            self._subs = [(p, c) for (p, c) in self._subs if not (p == pattern and c is client)]

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
    def publish(self, topic: str, payload, qos: int = 0, retain: bool = False) -> int:
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        payload = bytes(payload if payload is not None else b"")
        with self._lock:
            self._mid += 1
            mid = self._mid
            hooks = list(self._hooks)
            if retain:                          # empty retained payload clears the topic (MQTT-3.3.1-10)
                if payload:
                    self._retained[topic] = Message(topic, payload, qos, retain=True, mid=mid)
                else:
                    self._retained.pop(topic, None)
        deliveries: List[Tuple[float, Message]] = [(0.0, Message(topic, payload, qos, mid=mid))]
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
# TCP transport = MQTT 3.1.1 on the wire (control/mqtt.py): mosquitto_pub / mosquitto_sub / paho clients can talk
# to TcpBroker, and BusClient(transport="tcp") can talk to a real mosquitto.
# ---------------------------------------------------------------------------------------------
class _MqttBridgeClient:
    """Broker-side proxy for one MQTT connection; quacks like a BusClient for ``_enqueue``."""

    def __init__(self, wfile, client_id: str) -> None:
        self._wfile = wfile
        self._lock = threading.Lock()
        self.client_id = client_id
        self.alive = True
        self.taken_over = False

    def send(self, data: bytes) -> None:
        if not self.alive:
            return
        try:
            with self._lock:
                self._wfile.write(data)
                self._wfile.flush()
        except (OSError, ValueError):
            self.alive = False

    def _enqueue(self, msg: Message) -> None:
        # every subscription is granted qos 0, so deliveries never carry a packet id
        self.send(mqtt.publish(msg.topic, msg.payload, qos=0, retain=msg.retain))


class _MqttHandler(socketserver.StreamRequestHandler):
    CONNECT_TIMEOUT = 10.0
    MAX_PACKET = 1 << 20          # control-plane messages are tens of bytes; refuse to buffer more than 1 MiB

    def setup(self) -> None:
        # TLS handshake in the connection's own thread (a slow or hostile peer cannot stall the accept loop)
        ctx = self.server.owner.ssl_context  # type: ignore[attr-defined]
        self.tls_failed = False
        if ctx is not None:
            self.request.settimeout(self.CONNECT_TIMEOUT)
            try:
                self.request = ctx.wrap_socket(self.request, server_side=True)
            except (OSError, ValueError) as e:          # ssl.SSLError is an OSError
                log.info("TLS handshake with %s failed: %r", self.client_address, e)
                self.tls_failed = True
        super().setup()

    def handle(self) -> None:  # one thread per connection
        if self.tls_failed:
            return
        owner: "TcpBroker" = self.server.owner  # type: ignore[attr-defined]
        broker = owner.broker
        self.request.settimeout(self.CONNECT_TIMEOUT)
        try:
            pkt = mqtt.read_packet(self.rfile, self.MAX_PACKET)
            if pkt is None or pkt[0] != mqtt.CONNECT:
                return                                   # first packet must be CONNECT (MQTT-3.1.0-1)
            info = mqtt.parse_connect(pkt[2])
        except mqtt.ProtocolError:
            try:
                self.wfile.write(mqtt.connack(mqtt.CONNACK_BAD_PROTOCOL))
            except OSError:
                pass
            return
        except OSError:
            return
        client_id = info.client_id or owner.auto_client_id()
        proxy = _MqttBridgeClient(self.wfile, client_id)
        owner.register(client_id, proxy, self.request)
        proxy.send(mqtt.connack(mqtt.CONNACK_ACCEPTED))
        # a client that stays silent for 1.5 keep-alive periods is dead (MQTT-3.1.2-24) -> its will fires
        self.request.settimeout(1.5 * info.keepalive if info.keepalive else None)
        graceful = False
        try:
            while True:
                pkt = mqtt.read_packet(self.rfile, self.MAX_PACKET)
                if pkt is None:
                    break
                ptype, flags, body = pkt
                if ptype == mqtt.PUBLISH:
                    topic, payload, qos, retain, pid = mqtt.parse_publish(flags, body)
                    broker.publish(topic, payload, qos, retain=retain)
                    if qos == 1:
                        proxy.send(mqtt.puback(pid))
                    elif qos == 2:
                        proxy.send(mqtt.pubrec(pid))
                elif ptype == mqtt.PUBREL:
                    proxy.send(mqtt.pubcomp(mqtt.parse_packet_id(body)))
                elif ptype == mqtt.SUBSCRIBE:
                    pid, filters = mqtt.parse_subscribe(body)
                    proxy.send(mqtt.suback(pid, [0] * len(filters)))
                    for pattern, _qos in filters:
                        broker.add_subscription(pattern, proxy)  # type: ignore[arg-type]
                elif ptype == mqtt.UNSUBSCRIBE:
                    pid, filters = mqtt.parse_unsubscribe(body)
                    for pattern in filters:
                        broker.remove_subscription(pattern, proxy)  # type: ignore[arg-type]
                    proxy.send(mqtt.unsuback(pid))
                elif ptype == mqtt.PINGREQ:
                    proxy.send(mqtt.pingresp())
                elif ptype == mqtt.DISCONNECT:
                    graceful = True
                    break
                elif ptype in (mqtt.PUBACK, mqtt.PUBREC, mqtt.PUBCOMP):
                    pass
                else:
                    raise mqtt.ProtocolError(f"unexpected packet type {ptype}")
        except (OSError, mqtt.ProtocolError, ValueError):
            pass
        finally:
            proxy.alive = False
            broker.remove_client(proxy)  # type: ignore[arg-type]
            owner.unregister(client_id, proxy)
            if not graceful and not proxy.taken_over and info.will is not None:
                log.info("client %r vanished: publishing its will on %r", client_id, info.will.topic)
                broker.publish(info.will.topic, info.will.payload, info.will.qos, retain=info.will.retain)


class _ThreadedTCPServer(socketserver.ThreadingMixIn, socketserver.TCPServer):
    allow_reuse_address = True
    daemon_threads = True


class TcpBroker:
    """A small MQTT 3.1.1 broker on 127.0.0.1 (default port 1883) in front of an :class:`InProcessBroker` — the
    stand-in for mosquitto: qos-0 delivery, retained messages, last-will, keep-alive supervision, client-id
    take-over (a second connection with the same id closes the first, which is why workers use unique ids)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 1883,
                 broker: Optional[InProcessBroker] = None, ssl_context=None) -> None:
        self.broker = broker or InProcessBroker()
        self.ssl_context = ssl_context            # control/tls.py::server_context(...) -> MQTT over TLS (port 8883 by convention)
        self._server = _ThreadedTCPServer((host, port), _MqttHandler)
        self._server.owner = self  # type: ignore[attr-defined]
        self.host, self.port = self._server.server_address[:2]
        self._thread: Optional[threading.Thread] = None
        self._lock = threading.Lock()
        self._sessions: dict = {}     # client id -> (proxy, socket)
        self._auto = 0

    def auto_client_id(self) -> str:
        with self._lock:
            self._auto += 1
            return f"auto-{self._auto}"

    def register(self, client_id: str, proxy: _MqttBridgeClient, sock) -> None:
        with self._lock:
            old = self._sessions.get(client_id)
            self._sessions[client_id] = (proxy, sock)
        if old is not None:
            old[0].taken_over = True
            old[0].alive = False
            try:
                old[1].shutdown(socket.SHUT_RDWR)
            except OSError:
                pass

    def unregister(self, client_id: str, proxy: _MqttBridgeClient) -> None:
        with self._lock:
            cur = self._sessions.get(client_id)
            if cur is not None and cur[0] is proxy:
                del self._sessions[client_id]

    @property
    def clients(self) -> List[str]:
        with self._lock:
            return sorted(self._sessions)

    def start(self) -> "TcpBroker":
        self._thread = threading.Thread(target=self._server.serve_forever, name="bus-broker", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._server.shutdown()
        self._server.server_close()
        with self._lock:
            sessions = list(self._sessions.values())
        for proxy, sock in sessions:
            proxy.taken_over = True          # a broker going down is not a client failure: no wills
            try:
                sock.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
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
        self._will: Optional[mqtt.Will] = None
        self._ssl_context = None
        self._username: Optional[str] = None
        self._password: Optional[bytes] = None
        self._rfile = None
        self._keepalive = 0
        self._pinger: Optional[threading.Thread] = None
        self._closing = threading.Event()
        self._ack_cond = threading.Condition()
        self._acked: set = set()
        self._last_pong = 0.0
        # loop_start()/loop_forever() survive a raising callback (logged); drain()/loop() called directly from
        # tests re-raise unless this is set
        self.suppress_callback_errors = False
        # TCP transport: after an UNEXPECTED disconnect (broker restart, network blip) reconnect with exponential
        # back-off and restore the subscriptions — what paho's loop_forever() gives the reference for free.
        # reconnect_delay_set(0, 0) switches it off.
        self._reconnect_min, self._reconnect_max = 1.0, 30.0
        self._subscriptions: Dict[str, int] = {}
        self._addr: Optional[Tuple[str, int, int]] = None
        self._reconnector: Optional[threading.Thread] = None
        self.reconnects = 0

    # Default callbacks (paho signatures) — overridable --------------------------------
    def on_connect(self, client, userdata, flags, rc) -> None:  # noqa: D401
        pass

    def on_message(self, client, userdata, msg: Message) -> None:
        pass

    def on_publish(self, client, userdata, mid) -> None:
        pass

    def on_disconnect(self, client, userdata, rc) -> None:
        pass

    # paho-compatible session options (set before connect) -----------------------------------
    def will_set(self, topic: str, payload=None, qos: int = 0, retain: bool = False) -> None:
        """Last-will: published by the broker if this client vanishes without DISCONNECT (e.g. a worker that dies
        announces ``NOT_READY`` through its will)."""
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        self._will = mqtt.Will(topic, bytes(payload or b""), qos, retain)

    def tls_set(self, ca_certs: Optional[str] = None, certfile: Optional[str] = None, keyfile: Optional[str] = None,
                check_hostname: bool = False, context=None) -> None:
        """paho's ``tls_set``: MQTT over TLS (optionally with a client certificate for mutual TLS)."""
        from .tls import client_context
        self._ssl_context = context or client_context(ca_certs, certfile, keyfile, check_hostname=check_hostname)

    def reconnect_delay_set(self, min_delay: float = 1.0, max_delay: float = 30.0) -> None:
        """paho's ``reconnect_delay_set``; ``(0, 0)`` disables the automatic reconnect of the TCP transport."""
        self._reconnect_min, self._reconnect_max = float(min_delay), float(max_delay)

    def username_pw_set(self, username: Optional[str], password: Optional[str] = None) -> None:
        self._username = username
        self._password = password.encode("utf-8") if isinstance(password, str) else password

    # Connection ---------------------------------------------------------------------------
    def connect(self, host: str = "localhost", port: int = 1883, keepalive: int = 60) -> int:
        if self._transport == "tcp":
            addr = "127.0.0.1" if host in ("localhost", "") else host
            sock = socket.create_connection((addr, port), timeout=10)
            if self._ssl_context is not None:
                sock = self._ssl_context.wrap_socket(sock, server_hostname=addr)
            cid = self.client_id or f"colearn-{id(self):x}"
            sock.sendall(mqtt.connect(cid, keepalive, True, self._will, self._username, self._password))
            rfile = sock.makefile("rb")
            pkt = mqtt.read_packet(rfile)
            if pkt is None or pkt[0] != mqtt.CONNACK:
                sock.close()
                raise ConnectionError("broker closed the connection during the MQTT handshake")
            _, rc = mqtt.parse_connack(pkt[2])
            if rc != mqtt.CONNACK_ACCEPTED:
                sock.close()
                raise ConnectionRefusedError(f"MQTT broker refused the connection (return code {rc})")
            sock.settimeout(None)
            self._sock, self._rfile, self._keepalive = sock, rfile, keepalive
            self._addr = (host, port, keepalive)
            self._closing.clear()
            self._reader = threading.Thread(target=self._tcp_reader, name="bus-reader", daemon=True)
            self._reader.start()
            if keepalive > 0:
                self._pinger = threading.Thread(target=self._ping_loop, name="bus-ping", daemon=True)
                self._pinger.start()
        else:
            if self._broker is None:
                self._broker = default_broker()
        self._connected = True
        self.on_connect(self, None, {}, 0)
        return 0

    def _tcp_reader(self) -> None:
        rc = 1
        try:
            while True:
                pkt = mqtt.read_packet(self._rfile)
                if pkt is None:
                    break
                ptype, flags, body = pkt
                if ptype == mqtt.PUBLISH:
                    topic, payload, qos, retain, pid = mqtt.parse_publish(flags, body)
                    if qos == 1:
                        self._send(mqtt.puback(pid))
                    elif qos == 2:
                        self._send(mqtt.pubrec(pid))
                    self._enqueue(Message(topic, payload, qos, retain=retain, mid=pid))
                elif ptype == mqtt.PUBREL:
                    self._send(mqtt.pubcomp(mqtt.parse_packet_id(body)))
                elif ptype == mqtt.PUBREC:
                    self._send(mqtt.pubrel(mqtt.parse_packet_id(body)))
                elif ptype in (mqtt.SUBACK, mqtt.UNSUBACK, mqtt.PUBACK, mqtt.PUBCOMP):
                    with self._ack_cond:
                        self._acked.add((ptype, mqtt.parse_packet_id(body)))
                        self._ack_cond.notify_all()
                elif ptype == mqtt.PINGRESP:
                    self._last_pong = time.time()
        except (OSError, ValueError):
            pass
        finally:
            if self._closing.is_set():
                rc = 0
            self._connected = False
            try:
                self.on_disconnect(self, None, rc)
            except Exception:  # noqa: BLE001
                log.exception("on_disconnect raised")
            if rc != 0 and self._reconnect_max > 0 and self._addr is not None:
                self._reconnector = threading.Thread(target=self._reconnect_loop, name="bus-reconnect", daemon=True)
                self._reconnector.start()

    def _reconnect_loop(self) -> None:
        delay = max(0.05, self._reconnect_min)
        host, port, keepalive = self._addr  # type: ignore[misc]
        while not self._closing.is_set():
            if self._closing.wait(delay):
                return
            try:
                old = self._sock
                if old is not None:
                    try:
                        old.close()
                    except OSError:
                        pass
                self.connect(host, port, keepalive)
            except OSError as e:
                log.info("bus reconnect to %s:%d failed (%r); next attempt in %.1f s", host, port, e, min(delay * 2, self._reconnect_max))
                delay = min(delay * 2, self._reconnect_max)
                continue
            self.reconnects += 1
            for topic, qos in list(self._subscriptions.items()):
                try:
                    self.subscribe(topic, qos)
                except OSError:
                    break                                   # dropped again: the new reader thread schedules the next attempt
            log.info("bus reconnected to %s:%d (%d subscription(s) restored)", host, port, len(self._subscriptions))
            return

    def _ping_loop(self) -> None:
        period = max(0.05, self._keepalive / 2.0)
        while not self._closing.wait(period):
            try:
                self._send(mqtt.pingreq())
            except (OSError, AssertionError):
                return

    def _send(self, data: bytes) -> None:
        sock = self._sock
        if sock is None:
            raise OSError("not connected")
        with self._sock_lock:
            sock.sendall(data)

    def _next_pid(self) -> int:
        with self._sock_lock:
            self._mid = self._mid % 0xFFFF + 1
            return self._mid

    def wait_for_ack(self, ptype: int, pid: int, timeout: float = 5.0) -> bool:
        """Block until the broker acknowledged packet ``pid`` (SUBACK / UNSUBACK / PUBACK / PUBCOMP)."""
        deadline = time.time() + timeout
        with self._ack_cond:
            while (ptype, pid) not in self._acked:
                left = deadline - time.time()
                if left <= 0 or not self._connected:
                    return (ptype, pid) in self._acked
                self._ack_cond.wait(min(left, 0.1))
            self._acked.discard((ptype, pid))
            return True

    def subscribe(self, topic: str, qos: int = 0):
        if self._transport == "tcp":
            self._subscriptions[topic] = qos          # restored after an automatic reconnect
            pid = self._next_pid()
            self._send(mqtt.subscribe(pid, [(topic, qos)]))
            self.wait_for_ack(mqtt.SUBACK, pid)       # like mosquitto_sub: return once the subscription is live
            return (0, pid)
        assert self._broker is not None, "connect() first"
        self._broker.add_subscription(topic, self)
        return (0, 0)

    def unsubscribe(self, topic: str):
        if self._transport == "tcp":
            self._subscriptions.pop(topic, None)
            pid = self._next_pid()
            self._send(mqtt.unsubscribe(pid, [topic]))
            self.wait_for_ack(mqtt.UNSUBACK, pid)
            return (0, pid)
        assert self._broker is not None, "connect() first"
        self._broker.remove_subscription(topic, self)
        return (0, 0)

    def publish(self, topic: str, payload=b"", qos: int = 0, retain: bool = False):
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        payload = bytes(payload if payload is not None else b"")
        if self._transport == "tcp":
            mid = self._next_pid()
            self._send(mqtt.publish(topic, payload, qos, retain, packet_id=mid if qos else 0))
        else:
            assert self._broker is not None, "connect() first"
            mid = self._broker.publish(topic, payload, qos, retain=retain)
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
            self._closing.set()
            try:
                self._send(mqtt.disconnect())            # graceful: the broker discards the will
            except OSError:
                pass
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
