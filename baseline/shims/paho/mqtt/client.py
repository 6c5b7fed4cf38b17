"""``paho.mqtt.client`` stand-in: the ``Client`` base class the reference's Coordinator derives from.

No network: ``publish`` delivers straight to the ``on_message`` of every stand-in client subscribed to the topic in this
process, which is all the reference arm needs (events are injected the way ``mosquitto_pub`` would)."""
from __future__ import annotations

import threading
from typing import Dict, List

_SUBSCRIBERS: Dict[str, List["Client"]] = {}
_LOCK = threading.Lock()


class MQTTMessage:
    def __init__(self, topic: str, payload: bytes, qos: int = 0) -> None:
        self.topic, self.payload, self.qos, self.retain, self.mid = topic, payload, qos, False, 0


class Client:
    def __init__(self, client_id: str = "", clean_session: bool = True, userdata=None, **_kw) -> None:
        self._client_id, self._userdata = client_id, userdata
        self._stop = threading.Event()

    # connection management: nothing to connect to
    def connect(self, host: str = "localhost", port: int = 1883, keepalive: int = 60, **_kw) -> int:
        cb = getattr(self, "on_connect", None)
        if callable(cb):
            cb(self, self._userdata, {}, 0)
        return 0

    def disconnect(self) -> int:
        self._stop.set()
        return 0

    def subscribe(self, topic: str, qos: int = 0):
        with _LOCK:
            _SUBSCRIBERS.setdefault(topic, []).append(self)
        return (0, 1)

    def publish(self, topic: str, payload=None, qos: int = 0, retain: bool = False):
        data = payload if isinstance(payload, (bytes, bytearray)) else str(payload).encode()
        with _LOCK:
            targets = list(_SUBSCRIBERS.get(topic, ()))
        for c in targets:
            cb = getattr(c, "on_message", None)
            if callable(cb):
                cb(c, c._userdata, MQTTMessage(topic, bytes(data), qos))
        return (0, 1)

    def loop_forever(self, *_a, **_kw) -> None:
        self._stop.wait()

    def loop_start(self) -> None:
        pass

    def loop_stop(self, *_a, **_kw) -> None:
        self._stop.set()
