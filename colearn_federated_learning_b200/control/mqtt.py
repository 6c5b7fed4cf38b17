"""MQTT 3.1.1 wire codec (the subset a CoLearn deployment uses) — no third-party dependency.

The reference's control plane is a stock MQTT broker (mosquitto) with paho clients: the coordinator subscribes to
one topic, devices publish ``(ip[, port], STATE)`` with qos 0 (``federated_coordinator.py:92,294-300``,
``remote_worker.py:65-68,110``; README: ``mosquitto_pub -t topic/state -m "(192.168.1.7, TRAINING)"``).
Speaking the real wire protocol on the TCP transport means those tools keep working against
``federated_coordinator.py --embedded-broker`` and this repo's coordinator / workers can sit behind a real
mosquitto.

Implemented: CONNECT (3.1.1 ``MQTT``/4 and 3.1 ``MQIsdp``/3; clean session, will, username/password parsed),
CONNACK, PUBLISH qos 0/1/2 inbound (acknowledged) and qos 0 outbound, PUBACK / PUBREC / PUBREL / PUBCOMP,
SUBSCRIBE / SUBACK (every subscription is granted qos 0, which the spec allows), UNSUBSCRIBE / UNSUBACK,
PINGREQ / PINGRESP, DISCONNECT, retained messages and last-will.  Not implemented: persistent sessions, qos > 0
delivery to subscribers, MQTT 5.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import BinaryIO, Dict, List, Optional, Tuple

CONNECT, CONNACK, PUBLISH, PUBACK, PUBREC, PUBREL, PUBCOMP = 1, 2, 3, 4, 5, 6, 7
SUBSCRIBE, SUBACK, UNSUBSCRIBE, UNSUBACK, PINGREQ, PINGRESP, DISCONNECT = 8, 9, 10, 11, 12, 13, 14

CONNACK_ACCEPTED, CONNACK_BAD_PROTOCOL, CONNACK_ID_REJECTED = 0, 1, 2
MAX_PACKET = 256 * 1024 * 1024 - 1      # the protocol's own limit (4-byte variable length)


class ProtocolError(ValueError):
    """Malformed or unsupported packet; the connection must be closed (MQTT-4.8.0-1)."""


# ---------------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------------
def encode_remaining_length(n: int) -> bytes:
    if not 0 <= n <= MAX_PACKET:
        raise ProtocolError(f"remaining length {n} out of range")
    out = bytearray()
    while True:
        byte, n = n % 128, n // 128
        out.append(byte | (0x80 if n else 0))
        if not n:
            return bytes(out)


def pack_str(s) -> bytes:
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    if len(b) > 0xFFFF:
        raise ProtocolError("string longer than 65535 bytes")
    return struct.pack(">H", len(b)) + b


def packet(ptype: int, flags: int, body: bytes = b"") -> bytes:
    return bytes([(ptype << 4) | (flags & 0x0F)]) + encode_remaining_length(len(body)) + body


class _Cursor:
    def __init__(self, data: bytes) -> None:
        self.data, self.pos = data, 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise ProtocolError("packet shorter than its fields")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def u8(self) -> int:
        return self.take(1)[0]

    def u16(self) -> int:
        return struct.unpack(">H", self.take(2))[0]

    def string(self) -> bytes:
        return self.take(self.u16())

    def rest(self) -> bytes:
        out = self.data[self.pos:]
        self.pos = len(self.data)
        return out

    @property
    def done(self) -> bool:
        return self.pos >= len(self.data)


def _read_exact(f: BinaryIO, n: int) -> Optional[bytes]:
    buf = bytearray()
    while len(buf) < n:
        chunk = f.read(n - len(buf))
        if not chunk:
            return None
        buf.extend(chunk)
    return bytes(buf)


def read_packet(f: BinaryIO, max_size: int = MAX_PACKET) -> Optional[Tuple[int, int, bytes]]:
    """Read one control packet from a blocking binary stream: ``(type, flags, body)``; None on a clean EOF.
    ``max_size`` bounds the body a peer can make us allocate (the broker uses 1 MiB: events are tens of bytes)."""
    first = f.read(1)
    if not first:
        return None
    mult, length = 1, 0
    for i in range(4):
        b = f.read(1)
        if not b:
            return None
        length += (b[0] & 0x7F) * mult
        if not b[0] & 0x80:
            break
        mult *= 128
    else:
        raise ProtocolError("malformed remaining length")
    if length > max_size:
        raise ProtocolError(f"packet of {length} bytes exceeds the {max_size}-byte limit")
    body = _read_exact(f, length) if length else b""
    if body is None:
        return None
    return first[0] >> 4, first[0] & 0x0F, body


# ---------------------------------------------------------------------------------------------------------------
# packets
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Will:
    topic: str
    payload: bytes
    qos: int = 0
    retain: bool = False


@dataclass
class ConnectInfo:
    client_id: str
    keepalive: int
    clean_session: bool
    protocol_level: int
    will: Optional[Will] = None
    username: Optional[str] = None
    password: Optional[bytes] = None


def connect(client_id: str, keepalive: int = 60, clean_session: bool = True, will: Optional[Will] = None,
            username: Optional[str] = None, password: Optional[bytes] = None) -> bytes:
    flags = 0x02 if clean_session else 0
    payload = pack_str(client_id)
    if will is not None:
        flags |= 0x04 | ((will.qos & 3) << 3) | (0x20 if will.retain else 0)
        payload += pack_str(will.topic) + pack_str(will.payload)
    if username is not None:
        flags |= 0x80
        payload += pack_str(username)
        if password is not None:
            flags |= 0x40
            payload += pack_str(password)
    body = pack_str("MQTT") + bytes([4, flags]) + struct.pack(">H", keepalive & 0xFFFF) + payload
    return packet(CONNECT, 0, body)


def parse_connect(body: bytes) -> ConnectInfo:
    c = _Cursor(body)
    name, level = c.string(), c.u8()
    if (name, level) not in ((b"MQTT", 4), (b"MQIsdp", 3)):
        raise ProtocolError(f"unsupported protocol {name!r} level {level}")
    flags, keepalive = c.u8(), c.u16()
    if flags & 0x01:
        raise ProtocolError("reserved connect flag set")
    client_id = c.string().decode("utf-8", "replace")
    will = None
    if flags & 0x04:
        topic, payload = c.string().decode("utf-8", "replace"), c.string()
        will = Will(topic, payload, (flags >> 3) & 3, bool(flags & 0x20))
    username = c.string().decode("utf-8", "replace") if flags & 0x80 else None
    password = c.string() if flags & 0x40 else None
    return ConnectInfo(client_id, keepalive, bool(flags & 0x02), level, will, username, password)


def connack(return_code: int = CONNACK_ACCEPTED, session_present: bool = False) -> bytes:
    return packet(CONNACK, 0, bytes([1 if session_present else 0, return_code]))


def parse_connack(body: bytes) -> Tuple[bool, int]:
    if len(body) != 2:
        raise ProtocolError("CONNACK must be 2 bytes")
    return bool(body[0] & 1), body[1]


def publish(topic: str, payload: bytes = b"", qos: int = 0, retain: bool = False, dup: bool = False,
            packet_id: int = 0) -> bytes:
    if not topic or any(ch in topic for ch in "+#"):
        raise ProtocolError("a PUBLISH topic must be non-empty and free of wildcards")
    body = pack_str(topic)
    if qos:
        if not packet_id:
            raise ProtocolError("qos > 0 needs a packet id")
        body += struct.pack(">H", packet_id)
    return packet(PUBLISH, (0x08 if dup else 0) | ((qos & 3) << 1) | (1 if retain else 0), body + bytes(payload))


def parse_publish(flags: int, body: bytes) -> Tuple[str, bytes, int, bool, int]:
    """→ ``(topic, payload, qos, retain, packet_id)``"""
    qos = (flags >> 1) & 3
    if qos == 3:
        raise ProtocolError("qos 3 is not a thing")
    c = _Cursor(body)
    topic = c.string().decode("utf-8", "replace")
    pid = c.u16() if qos else 0
    return topic, c.rest(), qos, bool(flags & 1), pid


def _ack(ptype: int, packet_id: int) -> bytes:
    return packet(ptype, 0x02 if ptype == PUBREL else 0, struct.pack(">H", packet_id))


def puback(pid: int) -> bytes:
    return _ack(PUBACK, pid)


def pubrec(pid: int) -> bytes:
    return _ack(PUBREC, pid)


def pubrel(pid: int) -> bytes:
    return _ack(PUBREL, pid)


def pubcomp(pid: int) -> bytes:
    return _ack(PUBCOMP, pid)


def parse_packet_id(body: bytes) -> int:
    if len(body) < 2:
        raise ProtocolError("missing packet id")
    return struct.unpack(">H", body[:2])[0]


def subscribe(packet_id: int, filters: List[Tuple[str, int]]) -> bytes:
    body = struct.pack(">H", packet_id) + b"".join(pack_str(f) + bytes([q & 3]) for f, q in filters)
    return packet(SUBSCRIBE, 0x02, body)


def parse_subscribe(body: bytes) -> Tuple[int, List[Tuple[str, int]]]:
    c = _Cursor(body)
    pid, out = c.u16(), []
    while not c.done:
        out.append((c.string().decode("utf-8", "replace"), c.u8() & 3))
    if not out:
        raise ProtocolError("SUBSCRIBE without topic filters")
    return pid, out


def suback(packet_id: int, codes: List[int]) -> bytes:
    return packet(SUBACK, 0, struct.pack(">H", packet_id) + bytes(codes))


def unsubscribe(packet_id: int, filters: List[str]) -> bytes:
    return packet(UNSUBSCRIBE, 0x02, struct.pack(">H", packet_id) + b"".join(pack_str(f) for f in filters))


def parse_unsubscribe(body: bytes) -> Tuple[int, List[str]]:
    c = _Cursor(body)
    pid, out = c.u16(), []
    while not c.done:
        out.append(c.string().decode("utf-8", "replace"))
    return pid, out


def unsuback(packet_id: int) -> bytes:
    return _ack(UNSUBACK, packet_id)


def pingreq() -> bytes:
    return packet(PINGREQ, 0)


def pingresp() -> bytes:
    return packet(PINGRESP, 0)


def disconnect() -> bytes:
    return packet(DISCONNECT, 0)


PACKET_NAMES: Dict[int, str] = {CONNECT: "CONNECT", CONNACK: "CONNACK", PUBLISH: "PUBLISH", PUBACK: "PUBACK", PUBREC: "PUBREC",
                                PUBREL: "PUBREL", PUBCOMP: "PUBCOMP", SUBSCRIBE: "SUBSCRIBE", SUBACK: "SUBACK",
                                UNSUBSCRIBE: "UNSUBSCRIBE", UNSUBACK: "UNSUBACK", PINGREQ: "PINGREQ", PINGRESP: "PINGRESP",
                                DISCONNECT: "DISCONNECT"}
