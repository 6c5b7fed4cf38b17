"""Coordinator end-to-end on CPU: local VirtualWorker mode (BASELINE config 1), remote mode with real
worker servers, encrypted mode, inference, fault injection — all through the in-process bus + fake clock."""
import os
import socket
import threading
import time

import pytest
import torch

from colearn_federated_learning_b200 import settings
from colearn_federated_learning_b200.control.arguments import Arguments
from colearn_federated_learning_b200.control.bus import BusClient, InProcessBroker
from colearn_federated_learning_b200.control.coordinator import Coordinator, train_virtual_workers
from colearn_federated_learning_b200.control.window import FakeClock
from colearn_federated_learning_b200.control.workers import RemoteWorkerClient, WorkerServer
from colearn_federated_learning_b200.data import BaseDataset, synthetic_unsw, write_synthetic_csv, xor_toy_dataset
from colearn_federated_learning_b200.fl import FitConfig
from colearn_federated_learning_b200.models import FFNN, MLP, TestingRemote, flatten_params
from colearn_federated_learning_b200.utils.checkpoint import load_meta, save_model

TOPIC = "topic/state"
CPU = torch.device("cpu")


def make(tmp_path, remote=False, rounds=1, enc=False, iot=False, args=None, **kw):
    broker, clock = InProcessBroker(), FakeClock()
    a = args or Arguments()
    if not a.synthetic and not os.path.exists(a.test_path):
        a.synthetic = 64
    c = Coordinator(1, remote, rounds, enc, iot, args=a, broker=broker, timer_factory=clock,
                    path=str(tmp_path / "test.pth"), device=CPU, **kw)
    c.connect()
    c.subscribe(TOPIC)
    pub = BusClient("pub", broker=broker)
    pub.connect()
    return c, pub, clock


def test_local_mode_two_workers_one_round_baseline_config1(tmp_path):
    """BASELINE.json config 1: VirtualWorker mode, 2 workers on CPU, 10-feature MLP, 1 round."""
    csv = str(tmp_path / "train.csv")
    write_synthetic_csv(csv, 40, seed=1)
    args = Arguments(test_path=csv)
    c, pub, clock = make(tmp_path, args=args)
    pub.publish(TOPIC, "(192.168.1.7, TRAINING)")
    pub.publish(TOPIC, "(192.168.1.8, TRAINING)")
    assert c.drain() == 2 and len(settings.training_devices) == 2 and c.windower.state == "COLLECTING"
    clock.advance(1.0)
    res = c.windower.last_result
    assert res["workers"] == ["192.168.1.7", "192.168.1.8"] and set(res["losses"]) == set(res["workers"])
    assert os.path.exists(c.path) and len(settings.training_devices) == 0 and c.windower.state == "IDLE"
    state = torch.load(c.path, weights_only=True)
    assert list(state)[0] == "fc1.weight" and state["fc1.weight"].shape == (50, 10)
    assert load_meta(c.path)["mode"] == "local"


def test_local_mode_is_true_fedavg_not_the_alias_bug(tmp_path):
    torch.manual_seed(0)
    args = Arguments(synthetic=60, lr=0.05)
    c, pub, clock = make(tmp_path, args=args)
    init = FFNN()
    save_model(init, c.path)
    for ip in ("10.0.0.1", "10.0.0.2", "10.0.0.3"):
        pub.publish(TOPIC, f"({ip}, TRAINING)")
    c.drain()
    clock.advance(1.0)
    # recompute by hand: every worker starts from the same theta, uniform mean afterwards
    from colearn_federated_learning_b200.data import federate
    x, y = synthetic_unsw(60, seed=args.seed)
    fed = federate(BaseDataset(x, y), ["10.0.0.1", "10.0.0.2", "10.0.0.3"], CPU)
    theta = flatten_params(init).clone()
    flats, _, _ = train_virtual_workers(theta, FFNN(), fed, c._fit_config("local"), 0)
    got = torch.load(c.path, weights_only=True)
    m = FFNN()
    m.load_state_dict(got)
    assert torch.allclose(flatten_params(m), flats.mean(0), atol=1e-6)


def test_not_ready_and_invalid_events(tmp_path):
    c, pub, clock = make(tmp_path, args=Arguments(synthetic=32))
    for payload in ("(192.168.1.7, TRAINING)", "(192.168.1.8, TRAINING)", "(192.168.1.7, NOT_READY)",
                    "(garbage, TRAINING)", "(192.168.1.9, DANCING)", "nonsense", "(192.168.1.9)"):
        pub.publish(TOPIC, payload)
    c.drain()
    assert list(settings.training_devices.keys()) == ["192.168.1.8"]
    clock.advance(1.0)
    assert c.windower.last_result["workers"] == ["192.168.1.8"]


def test_iot_allow_list_filters_devices(tmp_path):
    f = tmp_path / "filtering_file.txt"
    f.write_text("192.168.1.7\n")
    c, pub, clock = make(tmp_path, iot=True, args=Arguments(synthetic=32), filter_file=str(f))
    pub.publish(TOPIC, "(192.168.1.7, TRAINING)")
    pub.publish(TOPIC, "(192.168.1.66, TRAINING)")
    c.drain()
    assert list(settings.training_devices.keys()) == ["192.168.1.7"]


def test_select_k_of_collected_workers(tmp_path):
    c, pub, clock = make(tmp_path, args=Arguments(synthetic=64, model="mlp"), select_k=2, selection="first")
    for i in range(4):
        pub.publish(TOPIC, f"(10.0.0.{i + 1}, TRAINING)")
    c.drain()
    clock.advance(1.0)
    assert c.windower.last_result["workers"] == ["10.0.0.1", "10.0.0.2"]
    assert list(settings.training_devices.keys()) == ["10.0.0.3", "10.0.0.4"]   # unselected stay for the next window


def test_rounds_reduce_loss_and_resume_from_checkpoint(tmp_path):
    args = Arguments(synthetic=256, model="mlp", lr=0.05, batch_size=4)
    c, pub, clock = make(tmp_path, rounds=4, args=args)
    pub.publish(TOPIC, "(10.0.0.1, TRAINING)")
    pub.publish(TOPIC, "(10.0.0.2, TRAINING)")
    c.drain()
    clock.advance(1.0)
    recs = c.metrics.records
    assert len(recs) == 4 and sum(recs[-1]["loss_k"]) < sum(recs[0]["loss_k"])
    first = torch.load(c.path, weights_only=True)["fc1.weight"].clone()
    pub.publish(TOPIC, "(10.0.0.1, TRAINING)")     # second window resumes from test.pth
    c.drain()
    clock.advance(1.0)
    assert not torch.equal(torch.load(c.path, weights_only=True)["fc1.weight"], first)
    assert c.trainings_done == 2


def test_encrypted_mode_two_workers(tmp_path, capsys):
    args = Arguments(synthetic=64, lr=0.1)
    args.n_train_items_enc = 12
    c, pub, clock = make(tmp_path, enc=True, args=args)
    pub.publish(TOPIC, "(10.0.0.1, TRAINING)")
    c.drain()
    clock.advance(1.0)
    assert c.windower.last_result is None                       # needs >= 2 workers (fc.py:114)
    for ip in ("10.0.0.2", "10.0.0.3", "10.0.0.4"):
        pub.publish(TOPIC, f"({ip}, TRAINING)")
    c.drain()
    clock.advance(1.0)
    res = c.windower.last_result
    assert len(res["workers"]) == 2 and res["batches"] == 12 and res["triples"] > 0
    assert os.path.exists(c.path)                                # the reference forgets to save (2.8-6)
    assert "Train Epoch: 0" in capsys.readouterr().out


# ---- remote mode with real worker servers ---------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(n=64, seed=0, inference=False, toy=False):
    port = _free_port()
    w = WorkerServer(f"127.0.0.1:{port}", "127.0.0.1", port, device=CPU)
    if toy:
        w.add_dataset(xor_toy_dataset())
    else:
        w.add_dataset(BaseDataset(*synthetic_unsw(n, seed=seed)))
    if inference:
        w.load_data([torch.rand(10) for _ in range(5)], tag="inference")
    w.start(block=False)
    return w, port


def test_remote_rounds_over_tcp_workers(tmp_path):
    w1, p1 = _worker(seed=1)
    w2, p2 = _worker(seed=2)
    try:
        c, pub, clock = make(tmp_path, remote=True, rounds=3, args=Arguments(lr=0.05), evaluate_after=False)
        pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {p2}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {_free_port()}, TRAINING)")     # nobody listening → skipped (fc.py:166-170)
        c.drain()
        assert len(settings.training_devices) == 2
        clock.advance(1.0)
        res = c.windower.last_result
        assert res["rounds"] == 3 and len(res["losses"]) == 2 and res["dropped"] == []
        assert c.args.federate_after_n_batches == 1000                     # round>1 rule (fc.py:533-535)
        assert w1.fits_served == 3 and w2.fits_served == 3
        assert os.path.exists(c.path) and len(settings.training_devices) == 0
    finally:
        w1.stop(); w2.stop()


def test_remote_worker_failure_is_dropped_not_fatal(tmp_path):
    good, pg = _worker(seed=1)
    bad, pb = _worker(toy=True)          # XOR toy data has 2 features → FFNN fit fails on that worker
    try:
        c, pub, clock = make(tmp_path, remote=True, rounds=2)
        pub.publish(TOPIC, f"(127.0.0.1, {pg}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {pb}, TRAINING)")
        c.drain()
        clock.advance(1.0)
        res = c.windower.last_result
        assert res["dropped"] == [f"127.0.0.1:{pb}"] and list(res["losses"]) == [f"127.0.0.1:{pg}"]
        assert good.fits_served == 2
    finally:
        good.stop(); bad.stop()


def test_inference_event_uses_trained_architecture(tmp_path):
    w, p = _worker(inference=True)
    try:
        c, pub, clock = make(tmp_path, remote=True)
        save_model(FFNN(), c.path)
        pub.publish(TOPIC, f"(127.0.0.1, {p}, INFERENCE)")
        c.drain()
        assert c.last_predictions is not None and len(c.last_predictions) == 5
        assert f"127.0.0.1:{p}" not in c.known_workers
    finally:
        w.stop()


def test_worker_rpc_direct():
    w, p = _worker(seed=3)
    try:
        cl = RemoteWorkerClient("x", "127.0.0.1", p)
        assert cl.ping() and cl.search("inference") == 0
        flat = flatten_params(FFNN())
        out, loss, n = cl.fit(flat, FitConfig(model="ffnn", loss="bce", max_nr_batches=10, lr=0.1))
        assert n == 64 and out.shape == flat.shape and not torch.equal(out, flat)
        with pytest.raises(RuntimeError):
            cl.fit(flat[:10], FitConfig(model="ffnn", loss="bce"))
        cl.close()
    finally:
        w.stop()
