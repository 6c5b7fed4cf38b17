"""Event grammar (reference event_parser.py) incl. the documented quirks and strict mode."""
import pytest
from hypothesis import given, strategies as st

from colearn_federated_learning_b200.control.event_parser import (EventParser, Event, STATES, format_event,
                                                                  valid_iot_ip_address)


def test_local_and_remote_forms():
    p = EventParser()
    p.set_message(b"(192.168.1.7, TRAINING)")
    assert p.ip_address() == "192.168.1.7" and p.state(local=True) == "TRAINING" and p.port(local=True) == -1
    p.set_message(b"(127.0.0.1, 8777, INFERENCE)")
    assert p.ip_address() == "127.0.0.1" and p.port() == 8777 and p.state() == "INFERENCE"
    assert p.training() == "TRAINING" and p.inference() == "INFERENCE"


def test_spaces_are_stripped_everywhere():
    p = EventParser()
    p.set_message(b" ( 10.0.0.1 ,  NOT_READY ) ")
    assert p.ip_address() == "10.0.0.1" and p.state(local=True) == "NOT_READY"


def test_invalid_state_and_ip():
    p = EventParser()
    p.set_message(b"(10.0.0.1, DANCING)")
    assert p.state(local=True) is None
    p.set_message(b"(not-an-ip, TRAINING)")
    assert p.ip_address() == -1


def test_reference_quirks_compat_vs_strict():
    lax, strict = EventParser(), EventParser(strict=True)
    for q in (lax, strict):
        q.set_message(b"(192.168.1.372, TRAINING)")
    assert lax.ip_address() == "192.168.1.372"      # no octet range check
    assert strict.ip_address() == -1
    for q in (lax, strict):
        q.set_message(b"(1.2.3.4junk, TRAINING)")
    assert lax.ip_address() == "1.2.3.4junk"         # prefix match
    assert strict.ip_address() == -1
    for q in (lax, strict):
        q.set_message(b"(1.2.3.4, 65535, TRAINING)")
    assert lax.port() == -1                           # range(65535) excludes 65535
    assert strict.port() == 65535


def test_short_tuple_raises_in_compat_but_not_in_strict_or_parse():
    lax = EventParser()
    lax.set_message(b"(1.2.3.4)")
    with pytest.raises(IndexError):
        lax.state(local=True)
    strict = EventParser(strict=True)
    strict.set_message(b"(1.2.3.4)")
    assert strict.state(local=True) is None and strict.port() == -1
    assert lax.parse(b"(1.2.3.4)", remote=False) is None
    assert lax.parse(b"\xff\xfe", remote=True) is None


def test_parse_and_worker_id():
    p = EventParser()
    assert p.parse("(10.0.0.2, TRAINING)", remote=False) == Event("10.0.0.2", "TRAINING")
    ev = p.parse(format_event("10.0.0.2", "TRAINING", 8778), remote=True)
    assert ev.worker_id == "10.0.0.2:8778" and ev.port == 8778


def test_allow_list(tmp_path):
    f = tmp_path / "filtering_file.txt"
    f.write_text("192.168.1.249\n192.168.1.22\n")
    assert valid_iot_ip_address("192.168.1.22", str(f))
    assert not valid_iot_ip_address("192.168.1.2", str(f))      # exact line match, no prefixes
    assert not valid_iot_ip_address("1.1.1.1", str(tmp_path / "missing.txt"))
    p = EventParser(filtering=True, filter_file=str(f))
    p.set_message(b"(192.168.1.249, TRAINING)")
    assert p.ip_address() == "192.168.1.249"
    p.set_message(b"(192.168.1.250, TRAINING)")
    assert p.ip_address() == -1
    f.write_text("192.168.1.250\n")                               # re-read on every event
    assert p.ip_address() == "192.168.1.250"


@given(st.tuples(*[st.integers(0, 255)] * 4), st.integers(1, 65534), st.sampled_from(STATES))
def test_roundtrip_property(octets, port, state):
    ip = ".".join(map(str, octets))
    for strict in (False, True):
        ev = EventParser(strict=strict).parse(format_event(ip, state, port), remote=True)
        assert ev == Event(ip, state, port)


@given(st.binary(max_size=64))
def test_parse_never_raises(payload):
    for strict in (False, True):
        for remote in (False, True):
            EventParser(strict=strict).parse(payload, remote=remote)
