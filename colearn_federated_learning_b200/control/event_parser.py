"""Event grammar of the signalling plane.

Wire format:  local ``"(<ipv4>, <STATE>)"``, remote ``"(<ipv4>, <port>, <STATE>)"``
with ``STATE ∈ {TRAINING, INFERENCE, NOT_READY}``.

Parity: reference ``event_parser.py:6-74`` (parser) and ``:76-91`` (allow-list scan).  Two modes:

* ``strict=False`` (default) reproduces the reference's observable behaviour, including its
  quirks: all spaces are stripped, the IPv4 check is a *prefix* regex match without octet range
  validation (so ``192.168.1.372`` and ``1.2.3.4junk`` pass — SURVEY §2.8-8), ports are valid in
  ``range(65535)`` (65535 itself is rejected).  Failures return ``-1`` / ``None`` like the
  reference.
* ``strict=True`` validates octets (0-255), requires a full match, and never raises on
  malformed payloads (the reference raises ``IndexError``/``ValueError`` on short tuples).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import List, Optional, Union

STATES = ["TRAINING", "INFERENCE", "NOT_READY"]
states = STATES  # reference-compatible alias (event_parser.py:3)

DEFAULT_FILTER_FILE = "./device_filtering/filtering_file.txt"

_IP_PREFIX = re.compile(r"\d{1,3}\.\d{1,3}\.\d{1,3}\.\d{1,3}")
_IP_FULL = re.compile(r"^(\d{1,3})\.(\d{1,3})\.(\d{1,3})\.(\d{1,3})$")


@dataclass(frozen=True)
class Event:
    """A parsed, validated event."""

    ip: str
    state: str
    port: Optional[int] = None

    @property
    def worker_id(self) -> str:
        # local id = ip (fc.py:149); remote id = "ip:port" (fc.py:162, rw.py:53)
        return self.ip if self.port is None else f"{self.ip}:{self.port}"


def valid_iot_ip_address(ip_address: str, filter_file: str = DEFAULT_FILTER_FILE) -> bool:
    """Exact-line match against the MUD-derived allow-list; the file is re-read on every event so
    that router-side updates (``file_upgrader.py``) take effect immediately (ep.py:76-91)."""
    try:
        with open(filter_file, "r") as f:
            for line in f:
                if line.rstrip() == ip_address:
                    return True
    except FileNotFoundError:
        return False
    return False


class EventParser:
    def __init__(self, filtering: bool = False, strict: bool = False,
                 filter_file: str = DEFAULT_FILTER_FILE) -> None:
        self._message = ""
        self.filtering = filtering
        self.strict = strict
        self.filter_file = filter_file

    # -- reference-compatible API -------------------------------------------------------
    def set_message(self, message: Union[bytes, str]) -> None:
        if isinstance(message, (bytes, bytearray)):
            message = bytes(message).decode("utf-8", errors="replace" if self.strict else "strict")
        self._message = message.replace(" ", "")

    def _fields(self) -> List[str]:
        return re.split(r",", re.sub(r"[\(\)]", "", self._message))

    def ip_address(self) -> Union[str, int]:
        ip = self._fields()[0]
        if self.strict:
            m = _IP_FULL.match(ip)
            ok = bool(m) and all(0 <= int(g) <= 255 for g in m.groups())
        else:
            ok = bool(_IP_PREFIX.match(ip))
        if not ok:
            return -1
        if self.filtering and not valid_iot_ip_address(ip, self.filter_file):
            return -1
        return ip

    def port(self, local: bool = False) -> int:
        if local:
            return -1
        fields = self._fields()
        try:
            port = int(fields[1])
        except (IndexError, ValueError):
            if self.strict:
                return -1
            raise
        upper = 65536 if self.strict else 65535  # reference: ``port in range(65535)``
        lower = 1 if self.strict else 0
        return port if lower <= port < upper else -1

    def state(self, local: bool = False) -> Optional[str]:
        fields = self._fields()
        idx = 1 if local else 2
        try:
            state = fields[idx]
        except IndexError:
            if self.strict:
                return None
            raise
        return state if state in STATES else None

    def training(self) -> str:
        return STATES[0]

    def inference(self) -> str:
        return STATES[1]

    def not_ready(self) -> str:
        return STATES[2]

    # -- convenience ------------------------------------------------------------------
    def parse(self, message: Union[bytes, str], remote: bool) -> Optional[Event]:
        """Parse + validate in one go; returns ``None`` for anything malformed (never raises)."""
        try:
            self.set_message(message)
            ip = self.ip_address()
            if ip == -1:
                return None
            state = self.state(local=not remote)
            if state is None:
                return None
            if remote:
                port = self.port()
                if port == -1:
                    return None
                return Event(ip=str(ip), state=state, port=port)
            return Event(ip=str(ip), state=state)
        except (IndexError, ValueError, UnicodeDecodeError):
            return None


def format_event(ip: str, state: str, port: Optional[int] = None) -> str:
    """Build the payload a device publishes (reference ``remote_worker.py:72``)."""
    if port is None:
        return f"({ip}, {state})"
    return f"({ip}, {port}, {state})"


def allowlist_path_from_env() -> str:
    return os.environ.get("COLEARN_FILTER_FILE", DEFAULT_FILTER_FILE)
