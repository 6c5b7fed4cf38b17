"""TLS on both TCP planes (control/tls.py): MQTT bus and worker RPC, incl. mutual TLS."""
import socket
import ssl
import time

import pytest
import torch

from colearn_federated_learning_b200.control import tls
from colearn_federated_learning_b200.control.bus import BusClient, TcpBroker
from colearn_federated_learning_b200.control.workers import RemoteWorkerClient, WorkerServer
from colearn_federated_learning_b200.data import BaseDataset, synthetic_unsw
from colearn_federated_learning_b200.fl import FitConfig
from colearn_federated_learning_b200.models import FFNN, flatten_params


@pytest.fixture(scope="module")
def pki(tmp_path_factory):
    return tls.make_test_pki(str(tmp_path_factory.mktemp("pki")))


def test_mqtt_over_tls_with_client_certificates(pki):
    server = tls.server_context(pki["server_cert"], pki["server_key"], pki["ca"], require_client_cert=True)
    with TcpBroker(port=0, ssl_context=server) as tb:
        got = []
        sub = BusClient("sub", transport="tcp")
        sub.tls_set(pki["ca"], pki["client_cert"], pki["client_key"], check_hostname=True)      # IP SAN 127.0.0.1
        sub.on_message = lambda c, u, m: got.append(m.payload)
        sub.connect(tb.host, tb.port)
        sub.subscribe("topic/state")
        pub = BusClient("pub", transport="tcp")
        pub.tls_set(pki["ca"], pki["client_cert"], pki["client_key"])
        pub.connect(tb.host, tb.port)
        pub.publish("topic/state", "(10.0.0.1, 8777, TRAINING)")
        assert sub.loop(5.0) == 1 and got == [b"(10.0.0.1, 8777, TRAINING)"]
        # no client certificate -> the handshake (or the first read) fails, the broker keeps serving
        anon = BusClient("anon", transport="tcp")
        anon.tls_set(pki["ca"])
        with pytest.raises((OSError, ConnectionError)):
            anon.connect(tb.host, tb.port)
        # plain TCP against the TLS port is refused as well
        plain = BusClient("plain", transport="tcp")
        with pytest.raises((OSError, ConnectionError)):
            plain.connect(tb.host, tb.port)
        pub.publish("topic/state", "still alive")
        assert sub.loop(5.0) == 1
        for c in (sub, pub):
            c.disconnect()


def test_worker_rpc_over_tls(pki):
    server = tls.server_context(pki["server_cert"], pki["server_key"], pki["ca"], require_client_cert=True)
    w = WorkerServer("w", "127.0.0.1", 0, device=torch.device("cpu"), ssl_context=server)
    w.add_dataset(BaseDataset(*synthetic_unsw(32, seed=1)))
    w.start(block=False)
    try:
        good = tls.client_context(pki["ca"], pki["client_cert"], pki["client_key"], check_hostname=True)
        cl = RemoteWorkerClient("w", "127.0.0.1", w.port, ssl_context=good)
        assert cl.ping()
        flat = flatten_params(FFNN())
        out, loss, n = cl.fit(flat, FitConfig(model="ffnn", loss="bce", max_nr_batches=4, lr=0.1))
        assert n == 32 and not torch.equal(out, flat)
        cl.close()
        # a coordinator without a certificate signed by the deployment's CA cannot drive the device
        with pytest.raises((OSError, ConnectionError, ssl.SSLError)):
            bad = RemoteWorkerClient("w", "127.0.0.1", w.port, ssl_context=tls.client_context(pki["ca"]), timeout=5)
            bad.ping()
        # wrong CA on the client side: the server is not trusted
        other = tls.make_test_pki(pki["ca"].rsplit("/", 1)[0] + "/other")
        with pytest.raises(ssl.SSLError):
            RemoteWorkerClient("w", "127.0.0.1", w.port, ssl_context=tls.client_context(other["ca"], pki["client_cert"], pki["client_key"]))
        # garbage on the TLS port does not take the server down
        s = socket.create_connection(("127.0.0.1", w.port), timeout=5)
        s.sendall(b"\x00" * 64)
        s.close()
        time.sleep(0.1)
        cl2 = RemoteWorkerClient("w", "127.0.0.1", w.port, ssl_context=good)
        assert cl2.ping()
        cl2.close()
    finally:
        w.stop()


def test_cli_helper(pki):
    assert tls.contexts_from_cli(None, None, None, server=False) is None
    with pytest.raises(SystemExit):
        tls.contexts_from_cli(None, "cert.pem", None, server=True)
    with pytest.raises(ValueError):
        tls.server_context(pki["server_cert"], pki["server_key"], None, require_client_cert=True)
    assert isinstance(tls.contexts_from_cli(pki["ca"], None, None, server=False), ssl.SSLContext)


def test_real_clis_over_mutual_tls(pki, tmp_path):
    """README remote flow with ``--tls-ca/--tls-cert/--tls-key`` on every process: the embedded broker and both
    devices demand certificates signed by the deployment's CA; one round trains and the checkpoint appears."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def free_port():
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s.close()
        return p

    bport, w1, w2 = free_port(), free_port(), free_port()
    ckpt = str(tmp_path / "test.pth")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=root)
    coord_tls = ["--tls-ca", pki["ca"], "--tls-cert", pki["server_cert"], "--tls-key", pki["server_key"]]
    dev_tls = ["--tls-ca", pki["ca"], "--tls-cert", pki["server_cert"], "--tls-key", pki["server_key"]]
    coord = subprocess.Popen([sys.executable, os.path.join(root, "federated_coordinator.py"), "-t", "topic/state", "-r", "-w", "2",
                              "-p", str(bport), "--host", "127.0.0.1", "--embedded-broker", "--checkpoint", ckpt, "--max-batches", "10",
                              "--exit-after", "1", "--no-cuda", *coord_tls],
                             env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    workers = []
    try:
        time.sleep(4.0)
        for port in (w1, w2):
            workers.append(subprocess.Popen([sys.executable, os.path.join(root, "remote_worker.py"), "--host", "127.0.0.1", "-p", str(port),
                                             "-b", "127.0.0.1", "--broker-port", str(bport), "-t", "topic/state", "-w", "1",
                                             "--synthetic", "32", "--no-cuda", *dev_tls], env=env,
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        out, _ = coord.communicate(timeout=120)
        assert coord.returncode == 0, out[-2000:]
        assert os.path.exists(ckpt) and out.count("Loss for worker id: 127.0.0.1:") == 2
    finally:
        for w in workers:
            w.kill()
        if coord.poll() is None:
            coord.kill()
