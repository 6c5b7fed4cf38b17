"""In-process pub/sub stub (incl. fault injection + TCP transport) and the temporal window."""
import threading
import time

import pytest
from hypothesis import given, strategies as st

from colearn_federated_learning_b200.control.bus import (BusClient, InProcessBroker, TcpBroker, topic_matches)
from colearn_federated_learning_b200.control.window import FakeClock, TemporalWindow
from colearn_federated_learning_b200.settings import DeviceRegistry


def test_topic_filters():
    assert topic_matches("topic/state", "topic/state")
    assert topic_matches("topic/+", "topic/state") and not topic_matches("topic/+", "topic/a/b")
    assert topic_matches("topic/#", "topic/a/b") and topic_matches("#", "x")
    assert not topic_matches("topic/state", "topic/other")


def _collector(broker, topic="topic/state"):
    got = []
    c = BusClient("sub", broker=broker)
    c.on_message = lambda cl, ud, m: got.append(m.payload)
    c.connect()
    c.subscribe(topic)
    return c, got


def test_publish_subscribe_and_callbacks():
    b = InProcessBroker()
    c, got = _collector(b)
    pubs = []
    p = BusClient("pub", broker=b)
    p.on_publish = lambda cl, ud, mid: pubs.append(mid)
    p.connect()
    p.publish("topic/state", "(1.2.3.4, TRAINING)")
    p.publish("topic/other", "nope")
    assert c.drain() == 1 and got == [b"(1.2.3.4, TRAINING)"] and len(pubs) == 2


def test_fault_injection_drop_duplicate_delay():
    b = InProcessBroker()
    c, got = _collector(b)
    p = BusClient("pub", broker=b)
    p.connect()
    b.inject_drop(lambda m: b"DROP" in m.payload)
    b.inject_duplicate(lambda m: b"DUP" in m.payload)
    p.publish("topic/state", "DROP me")
    p.publish("topic/state", "DUP me")
    c.drain()
    assert got == [b"DUP me", b"DUP me"] and b.dropped == 1
    b.clear_fault_hooks()
    b.inject_delay(lambda m: True, 0.05)
    p.publish("topic/state", "late")
    assert c.drain() == 0
    time.sleep(0.15)
    assert c.drain() == 1 and got[-1] == b"late"


def test_fault_injection_from_cli_spec():
    b = InProcessBroker()
    c, got = _collector(b)
    p = BusClient("pub", broker=b)
    p.connect()
    b.inject_from_spec(r"drop:192\.168\.1\.66")
    b.inject_from_spec("dup:NOT_READY")
    p.publish("topic/state", "(192.168.1.66, TRAINING)")
    p.publish("topic/state", "(192.168.1.7, NOT_READY)")
    p.publish("topic/state", "(192.168.1.8, TRAINING)")
    c.drain()
    assert got == [b"(192.168.1.7, NOT_READY)"] * 2 + [b"(192.168.1.8, TRAINING)"]
    with pytest.raises(ValueError):
        b.inject_from_spec("explode:everything")


def test_tcp_transport_roundtrip():
    with TcpBroker("127.0.0.1", 0) as broker:
        got = []
        ev = threading.Event()
        sub = BusClient("s", transport="tcp")
        sub.on_message = lambda cl, ud, m: (got.append((m.topic, m.payload)), ev.set())
        sub.connect("127.0.0.1", broker.port)
        sub.subscribe("topic/#")
        sub.loop_start()
        time.sleep(0.05)
        pub = BusClient("p", transport="tcp")
        pub.connect("127.0.0.1", broker.port)
        pub.publish("topic/state", b"(127.0.0.1, 8777, TRAINING)")
        assert ev.wait(5)
        assert got == [("topic/state", b"(127.0.0.1, 8777, TRAINING)")]
        sub.disconnect()
        pub.disconnect()


# ---------------------------------------------------------------------------------------------
def _window(window=2.0, **kw):
    reg, clock, calls = DeviceRegistry(), FakeClock(), []

    def train(snapshot):
        calls.append(list(snapshot))
        for wid in snapshot:
            reg.remove(wid)
        return len(snapshot)

    return reg, clock, calls, TemporalWindow(reg, window, train, timer_factory=clock, **kw)


def test_first_event_arms_later_events_join_the_same_window():
    reg, clock, calls, w = _window()
    assert w.on_training("a", "A") is True and w.state == w.COLLECTING
    clock.advance(1.0)
    assert w.on_training("b", "B") is False
    clock.advance(0.99)
    assert calls == []
    clock.advance(0.02)
    assert calls == [["a", "b"]] and w.state == w.IDLE and reg.event_served == 0


def test_not_ready_inside_window_removes_device():
    reg, clock, calls, w = _window()
    w.on_training("a", "A")
    w.on_training("b", "B")
    w.on_not_ready("a")
    clock.advance(2.0)
    assert calls == [["b"]]


def test_late_joiner_rides_next_window_keep_policy():
    reg, clock, calls = DeviceRegistry(), FakeClock(), []
    holder = {}

    def train(snapshot):
        calls.append(list(snapshot))
        holder["w"].on_training("late", "L")   # arrives while training runs
        for wid in snapshot:
            reg.remove(wid)

    w = TemporalWindow(reg, 1.0, train, timer_factory=clock)
    holder["w"] = w
    w.on_training("a", "A")
    clock.advance(1.0)
    assert calls == [["a"]] and "late" in reg and w.state == w.IDLE and clock.pending == 0
    w.on_training("b", "B")                    # next TRAINING event arms the next window
    clock.advance(1.0)
    assert calls == [["a"], ["late", "b"]]


def test_rearm_if_pending_option():
    reg, clock, calls = DeviceRegistry(), FakeClock(), []
    holder = {}

    def train(snapshot):
        calls.append(list(snapshot))
        if len(calls) == 1:
            holder["w"].on_training("late", "L")
        for wid in snapshot:
            reg.remove(wid)

    w = TemporalWindow(reg, 1.0, train, timer_factory=clock, rearm_if_pending=True)
    holder["w"] = w
    w.on_training("a", "A")
    clock.advance(1.0)
    clock.advance(1.0)
    assert calls == [["a"], ["late"]]


def test_crashing_trainer_never_wedges_the_window():
    reg, clock = DeviceRegistry(), FakeClock()

    def boom(snapshot):
        raise RuntimeError("trainer died")

    w = TemporalWindow(reg, 1.0, boom, timer_factory=clock)
    w.on_training("a", "A")
    clock.advance(1.0)
    assert isinstance(w.last_error, RuntimeError) and w.state == w.IDLE and reg.event_served == 0
    assert w.on_training("b", "B") is True


def test_below_lower_bound_resets_without_training():
    reg, clock, calls, w = _window(lower_bound=2)
    w.on_training("a", "A")
    clock.advance(2.0)
    assert calls == [] and w.history[-1]["trained"] is False and reg.event_served == 0


def test_cancel_only_while_collecting():
    reg, clock, calls, w = _window()
    assert w.cancel() is False
    w.on_training("a", "A")
    assert w.cancel() is True
    clock.advance(5.0)
    assert calls == []


@given(st.lists(st.tuples(st.sampled_from(["T", "N", "tick"]), st.integers(0, 5)), max_size=40))
def test_window_invariants(ops_):
    reg, clock, calls, w = _window(window=1.0)
    for op, k in ops_:
        if op == "T":
            w.on_training(f"d{k}", k)
        elif op == "N":
            w.on_not_ready(f"d{k}")
        else:
            clock.advance(0.4 * (k + 1))
        assert clock.pending <= 1                       # never two windows armed
        assert (w.state == w.COLLECTING) == (clock.pending == 1)
    for c in calls:
        assert len(set(c)) == len(c)


def test_registry_is_thread_safe():
    reg = DeviceRegistry()

    def work(i):
        for j in range(200):
            reg.register(f"{i}-{j}", j)
            reg.serve_event()
            reg.snapshot()
            reg.remove(f"{i}-{j}")

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(reg) == 0 and reg.event_served == 1600


def test_network_loop_survives_a_raising_callback_and_malformed_tcp_frames():
    """A bug in ``on_message`` (or a garbage frame on the TCP transport) must not silently kill the control plane."""
    import socket
    import time

    from colearn_federated_learning_b200.control.bus import BusClient, InProcessBroker, TcpBroker

    broker = InProcessBroker()
    seen = []

    class Flaky(BusClient):
        def on_message(self, client, userdata, msg):
            if msg.payload == b"boom":
                raise RuntimeError("handler bug")
            seen.append(msg.payload)

    c = Flaky("flaky", broker=broker)
    c.connect()
    c.subscribe("t")
    c.loop_start()
    broker.publish("t", b"boom")
    broker.publish("t", b"ok")
    deadline = time.time() + 5
    while not seen and time.time() < deadline:
        time.sleep(0.01)
    c.loop_stop()
    assert seen == [b"ok"]
    # direct drain() (how the unit tests drive the bus) still surfaces the error
    broker.publish("t", b"boom")
    with pytest.raises(RuntimeError):
        c.drain()
    # delayed deliveries do not accumulate finished timers
    broker.inject_delay(lambda m: True, 0.01)
    for _ in range(20):
        broker.publish("t", b"x")
        time.sleep(0.002)
    time.sleep(0.1)
    broker.publish("t", b"x")
    assert len(broker._timers) <= 2

    with TcpBroker(port=0) as tb:
        # garbage instead of an MQTT CONNECT: that connection is dropped, the broker keeps serving
        s = socket.create_connection((tb.host, tb.port), timeout=5)
        s.sendall(b'{"op":"pub","topic":"t"}\n' + bytes(range(256)))
        s.settimeout(5)
        assert s.recv(16) in (b"", b"\x20\x02\x00\x01")           # closed (or CONNACK "bad protocol" then closed)
        s.close()
        got = []
        cli = BusClient("tcp", transport="tcp")
        cli.on_message = lambda c_, u, m: got.append(m.payload)
        cli.connect(tb.host, tb.port)
        cli.subscribe("t")
        pub = BusClient("tcp-pub", transport="tcp")
        pub.connect(tb.host, tb.port)
        pub.publish("t", b"hi")
        deadline = time.time() + 5
        while not got and time.time() < deadline:
            cli.loop(0.05)
        pub.disconnect()
        cli.disconnect()
        assert got == [b"hi"]
