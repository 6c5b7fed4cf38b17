"""MQTT 3.1.1 on the TCP transport (control/mqtt.py, TcpBroker, BusClient(transport="tcp")).

The byte strings below are written out by hand from the MQTT 3.1.1 specification (OASIS, §3): they are what
``mosquitto_pub`` / paho put on the wire, so a broker that answers them correctly interoperates with the tools the
reference's README tells its users to run (``mosquitto_pub -t topic/state -m "(192.168.1.7, TRAINING)"``).
"""
import io
import socket
import time

import pytest
from hypothesis import given, strategies as st

from colearn_federated_learning_b200.control import mqtt
from colearn_federated_learning_b200.control.bus import BusClient, InProcessBroker, TcpBroker

CONNECT_A = bytes.fromhex("100d" "00044d515454" "04" "02" "003c" "0001" "61")      # client id "a", clean, keepalive 60
CONNACK_OK = bytes.fromhex("20020000")
PUBLISH_AB_HI = bytes.fromhex("3007" "0003612f62" "6869")                           # qos 0, topic a/b, payload "hi"
SUBSCRIBE_AB = bytes.fromhex("8208" "0001" "0003612f62" "00")                       # packet id 1, a/b, qos 0
SUBACK_1 = bytes.fromhex("9003000100")
PINGREQ, PINGRESP, DISCONNECT = bytes.fromhex("c000"), bytes.fromhex("d000"), bytes.fromhex("e000")


def test_codec_matches_the_specification_byte_for_byte():
    assert mqtt.connect("a", 60) == CONNECT_A
    assert mqtt.connack() == CONNACK_OK
    assert mqtt.publish("a/b", b"hi") == PUBLISH_AB_HI
    assert mqtt.subscribe(1, [("a/b", 0)]) == SUBSCRIBE_AB
    assert mqtt.suback(1, [0]) == SUBACK_1
    assert (mqtt.pingreq(), mqtt.pingresp(), mqtt.disconnect()) == (PINGREQ, PINGRESP, DISCONNECT)
    assert mqtt.puback(0x1234) == bytes.fromhex("40021234") and mqtt.pubrel(7) == bytes.fromhex("62020007")
    assert mqtt.publish("t", b"x", qos=1, packet_id=10, retain=True) == bytes.fromhex("3306" "000174" "000a" "78")
    # remaining-length varint boundaries (spec table 2.4)
    for n, enc in ((0, "00"), (127, "7f"), (128, "8001"), (16383, "ff7f"), (16384, "808001"), (2097151, "ffff7f"),
                   (2097152, "80808001"), (268435455, "ffffff7f")):
        assert mqtt.encode_remaining_length(n) == bytes.fromhex(enc)
    with pytest.raises(mqtt.ProtocolError):
        mqtt.encode_remaining_length(268435456)
    # parsing
    info = mqtt.parse_connect(CONNECT_A[2:])
    assert (info.client_id, info.keepalive, info.clean_session, info.protocol_level, info.will) == ("a", 60, True, 4, None)
    full = mqtt.connect("w1", 5, True, mqtt.Will("topic/state", b"(10.0.0.1, 8777, NOT_READY)", 1, True), "user", b"pw")
    t, f, body = mqtt.read_packet(io.BytesIO(full))
    info = mqtt.parse_connect(body)
    assert t == mqtt.CONNECT and info.will == mqtt.Will("topic/state", b"(10.0.0.1, 8777, NOT_READY)", 1, True)
    assert (info.username, info.password) == ("user", b"pw")
    legacy = mqtt.packet(mqtt.CONNECT, 0, mqtt.pack_str("MQIsdp") + bytes([3, 2, 0, 30]) + mqtt.pack_str("old"))
    assert mqtt.parse_connect(mqtt.read_packet(io.BytesIO(legacy))[2]).protocol_level == 3    # MQTT 3.1 clients
    with pytest.raises(mqtt.ProtocolError):
        mqtt.parse_connect(mqtt.pack_str("MQTT") + bytes([5, 2, 0, 30]) + mqtt.pack_str("v5"))
    assert mqtt.parse_publish(0, PUBLISH_AB_HI[2:]) == ("a/b", b"hi", 0, False, 0)
    assert mqtt.parse_subscribe(SUBSCRIBE_AB[2:]) == (1, [("a/b", 0)])
    with pytest.raises(mqtt.ProtocolError):
        mqtt.publish("a/+", b"")                     # wildcards are for filters only
    with pytest.raises(mqtt.ProtocolError):
        mqtt.parse_publish(0x06, PUBLISH_AB_HI[2:])  # qos 3


@given(topic=st.text(st.characters(blacklist_characters="+#\x00", blacklist_categories=("Cs",)), min_size=1, max_size=40),
       payload=st.binary(max_size=300), qos=st.integers(0, 2), retain=st.booleans())
def test_publish_roundtrip(topic, payload, qos, retain):
    raw = mqtt.publish(topic, payload, qos, retain, packet_id=77 if qos else 0)
    t, flags, body = mqtt.read_packet(io.BytesIO(raw + b"trailing"))
    assert t == mqtt.PUBLISH and mqtt.parse_publish(flags, body) == (topic, payload, qos, retain, 77 if qos else 0)


def _recv(sock, n, timeout=5.0):
    sock.settimeout(timeout)
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            break
        buf += chunk
    return buf


def test_broker_speaks_to_a_raw_mosquitto_style_client():
    """What ``mosquitto_sub -t a/b`` and ``mosquitto_pub -t a/b -m hi`` send, byte for byte."""
    with TcpBroker(port=0) as tb:
        sub = socket.create_connection((tb.host, tb.port))
        sub.sendall(CONNECT_A)
        assert _recv(sub, 4) == CONNACK_OK
        sub.sendall(SUBSCRIBE_AB)
        assert _recv(sub, 5) == SUBACK_1
        sub.sendall(PINGREQ)
        assert _recv(sub, 2) == PINGRESP
        pub = socket.create_connection((tb.host, tb.port))
        pub.sendall(mqtt.connect("b", 60) + PUBLISH_AB_HI + DISCONNECT)      # fire and forget, like mosquitto_pub
        assert _recv(pub, 4) == CONNACK_OK
        assert _recv(sub, len(PUBLISH_AB_HI)) == PUBLISH_AB_HI
        # qos 1 and qos 2 publishes are acknowledged and delivered at the granted qos 0
        pub2 = socket.create_connection((tb.host, tb.port))
        pub2.sendall(mqtt.connect("c", 60) + mqtt.publish("a/b", b"q1", qos=1, packet_id=5))
        assert _recv(pub2, 4) == CONNACK_OK and _recv(pub2, 4) == mqtt.puback(5)
        assert _recv(sub, 9) == mqtt.publish("a/b", b"q1")
        pub2.sendall(mqtt.publish("a/b", b"q2", qos=2, packet_id=6))
        assert _recv(pub2, 4) == mqtt.pubrec(6)
        pub2.sendall(mqtt.pubrel(6))
        assert _recv(pub2, 4) == mqtt.pubcomp(6)
        assert _recv(sub, 9) == mqtt.publish("a/b", b"q2")
        # unsubscribe stops the flow
        sub.sendall(mqtt.unsubscribe(2, ["a/b"]))
        assert _recv(sub, 4) == mqtt.unsuback(2)
        pub2.sendall(PUBLISH_AB_HI + PINGREQ)
        assert _recv(pub2, 2) == PINGRESP
        sub.sendall(PINGREQ)
        assert _recv(sub, 2) == PINGRESP                 # nothing was queued in front of the ping response
        for s in (sub, pub, pub2):
            s.close()
        # not MQTT / wrong protocol level
        bad = socket.create_connection((tb.host, tb.port))
        bad.sendall(mqtt.packet(mqtt.CONNECT, 0, mqtt.pack_str("MQTT") + bytes([5, 2, 0, 30]) + mqtt.pack_str("v5")))
        assert _recv(bad, 4) == mqtt.connack(mqtt.CONNACK_BAD_PROTOCOL) and _recv(bad, 1) == b""
        bad.close()


def test_will_retained_takeover_and_keepalive():
    events = []
    with TcpBroker(port=0) as tb:
        watcher = BusClient("watcher", transport="tcp")
        watcher.on_message = lambda c, u, m: events.append((m.topic, m.payload, m.retain))
        watcher.connect(tb.host, tb.port)
        watcher.subscribe("topic/#")
        watcher.loop_start()

        def wait_for(n):
            deadline = time.time() + 10
            while len(events) < n and time.time() < deadline:
                time.sleep(0.01)
            return len(events) >= n

        # 1. a worker with a NOT_READY will dies without DISCONNECT -> the broker publishes the will
        w = BusClient("worker-1", transport="tcp")
        w.will_set("topic/state", "(10.0.0.1, 8777, NOT_READY)")
        w.connect(tb.host, tb.port, keepalive=0)
        w.publish("topic/state", "(10.0.0.1, 8777, TRAINING)")
        assert wait_for(1)
        w._sock.shutdown(socket.SHUT_RDWR)                # crash: the TCP connection drops without DISCONNECT
        assert wait_for(2) and events[1][1] == b"(10.0.0.1, 8777, NOT_READY)"
        # 2. graceful disconnect: no will
        w2 = BusClient("worker-2", transport="tcp")
        w2.will_set("topic/state", "(10.0.0.2, 8778, NOT_READY)")
        w2.connect(tb.host, tb.port)
        w2.disconnect()
        time.sleep(0.2)
        assert len(events) == 2
        # 3. keep-alive: a client that goes silent for 1.5 x keepalive is declared dead (will fires)
        silent = socket.create_connection((tb.host, tb.port))
        silent.sendall(mqtt.connect("silent", 1, True, mqtt.Will("topic/state", b"silent-died")))
        assert _recv(silent, 4) == CONNACK_OK
        t0 = time.time()
        assert wait_for(3) and events[2][1] == b"silent-died" and 1.0 < time.time() - t0 < 5.0
        silent.close()
        # ... while a BusClient keeps itself alive with PINGREQs
        alive = BusClient("alive", transport="tcp")
        alive.will_set("topic/state", "alive-died")
        alive.connect(tb.host, tb.port, keepalive=2)     # pings every second, the broker would give up after 3 s
        time.sleep(3.5)
        assert len(events) == 3 and "alive" in tb.clients
        # 4. take-over: a second connection with the same client id closes the first, without firing its will
        dis = []
        alive.on_disconnect = lambda c, u, rc: dis.append(rc)
        twin = BusClient("alive", transport="tcp")
        twin.connect(tb.host, tb.port)
        deadline = time.time() + 5
        while not dis and time.time() < deadline:
            time.sleep(0.01)
        assert dis == [1] and len(events) == 3
        twin.disconnect()
        # 5. retained message: delivered to later subscribers with the retain flag, cleared by an empty payload
        p = BusClient("p", transport="tcp")
        p.connect(tb.host, tb.port)
        p.publish("topic/config", b"window=30", retain=True)
        assert wait_for(4) and events[3] == ("topic/config", b"window=30", False)
        late = BusClient("late", transport="tcp")
        got = []
        late.on_message = lambda c, u, m: got.append((m.payload, m.retain))
        late.connect(tb.host, tb.port)
        late.subscribe("topic/config")
        assert late.loop(2.0) == 1 and got == [(b"window=30", True)]
        p.publish("topic/config", b"", retain=True)
        late2 = BusClient("late2", transport="tcp")
        late2.connect(tb.host, tb.port)
        late2.subscribe("topic/config")
        assert late2.loop(0.3) == 0
        for c in (p, late, late2):
            c.disconnect()
        watcher.disconnect()


def test_in_process_bus_has_the_same_retained_and_unsubscribe_semantics():
    b = InProcessBroker()
    got = []
    c = BusClient("c", broker=b)
    c.on_message = lambda cl, u, m: got.append((m.topic, m.payload, m.retain))
    c.connect()
    b.publish("cfg/x", b"1", retain=True)
    c.subscribe("cfg/+")
    c.subscribe("cfg/+")                                  # re-subscribing re-sends the retained message (MQTT-3.8.4-3)...
    b.publish("cfg/x", b"2")                              # ...but does not duplicate the subscription
    assert c.drain() == 3 and got == [("cfg/x", b"1", True), ("cfg/x", b"1", True), ("cfg/x", b"2", False)]
    c.unsubscribe("cfg/+")
    b.publish("cfg/x", b"3")
    assert c.drain() == 0


def test_client_reconnects_after_a_broker_restart_and_restores_subscriptions():
    """paho's loop_forever() reconnects after a broker outage; so does the TCP bus client: exponential back-off, then the
    subscriptions are restored and messages flow again.  A deliberate disconnect() does not reconnect."""
    import time

    from colearn_federated_learning_b200.control.bus import BusClient, TcpBroker

    broker = TcpBroker("127.0.0.1", 0).start()
    port = broker.port
    got = []
    sub = BusClient("sub", transport="tcp")
    sub.reconnect_delay_set(0.1, 0.5)
    sub.on_message = lambda c, u, m: got.append(m.payload)
    down = []
    sub.on_disconnect = lambda c, u, rc: down.append(rc)
    sub.connect("127.0.0.1", port)
    sub.subscribe("topic/state")
    sub.loop_start()
    pub = BusClient("pub", transport="tcp")
    pub.connect("127.0.0.1", port)
    pub.publish("topic/state", "one")
    deadline = time.time() + 5
    while not got and time.time() < deadline:
        time.sleep(0.01)
    assert got == [b"one"]
    pub.disconnect()
    broker.stop()                                            # outage: the subscriber loses its connection
    deadline = time.time() + 5
    while not down and time.time() < deadline:
        time.sleep(0.01)
    assert down and down[0] != 0
    time.sleep(0.4)                                          # a few failed attempts while nothing listens
    broker2 = TcpBroker("127.0.0.1", port).start()
    try:
        deadline = time.time() + 10
        while sub.reconnects == 0 and time.time() < deadline:
            time.sleep(0.02)
        assert sub.reconnects == 1
        pub2 = BusClient("pub2", transport="tcp")
        pub2.connect("127.0.0.1", port)
        deadline = time.time() + 5
        while len(got) < 2 and time.time() < deadline:
            pub2.publish("topic/state", "two")               # the restored subscription delivers again
            time.sleep(0.05)
        assert got[-1] == b"two"
        sub.disconnect()                                     # deliberate: rc 0, no reconnect
        time.sleep(0.3)
        assert sub.reconnects == 1 and down[-1] == 0
        pub2.disconnect()
    finally:
        broker2.stop()
