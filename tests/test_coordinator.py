"""Coordinator end-to-end on CPU: local VirtualWorker mode (BASELINE config 1), remote mode with real
worker servers, encrypted mode, inference, fault injection — all through the in-process bus + fake clock."""
import os
import socket
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
from colearn_federated_learning_b200.models import FFNN, flatten_params
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
        # traffic accounting (paper §4.3): per round and worker one fit request (2 401 fp32 parameters + config) out and
        # the trained parameters + loss back — ~9.7 kB each way against the reference's ~36 kB / ~30.6 kB
        recs = [r for r in c.metrics.records if "bytes_out" in r]
        assert len(recs) == 3 and all(2 * 9604 < r["bytes_out"] < 2 * 11000 and 2 * 9604 < r["bytes_in"] < 2 * 11000 for r in recs)
        assert res["bytes_out"] == sum(r["bytes_out"] for r in recs) and res["bytes_in"] == sum(r["bytes_in"] for r in recs)
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


def test_worker_runtime_reuses_model_and_arena_and_guards_the_wire():
    """Robustness of the RPC runtime: one model/arena per architecture across fits (trainer caches key on them),
    a handle that timed out is dead (no desynchronised replies), oversized frames are refused."""
    import socket
    import struct

    from colearn_federated_learning_b200.control import workers as W

    w, p = _worker(seed=5)
    try:
        cl = RemoteWorkerClient("x", "127.0.0.1", p)
        flat = flatten_params(FFNN())
        cfg = FitConfig(model="ffnn", loss="bce", max_nr_batches=5, lr=0.1)
        a, _, _ = cl.fit(flat, cfg)
        model0, arena0 = w._models["ffnn"]
        b, _, _ = cl.fit(flat, cfg)
        assert w._models["ffnn"][0] is model0 and w._models["ffnn"][1].data_ptr() == arena0.data_ptr()
        assert w.fits_served == 2 and a.shape == b.shape
        # the arena is overwritten by every request: fit #2 started from `flat`, not from fit #1's result
        c, _, _ = cl.fit(a, cfg)
        assert not torch.equal(c, b)
        # timeout -> the handle closes itself
        with pytest.raises(OSError):
            # 2 000 epochs x 64 batch-1 steps: ~0.3 s even through the native host executor (~2 us / step)
            cl._call({"op": "fit", "config": FitConfig(model="ffnn", loss="bce", epochs=2000).to_dict(),
                      "params": W._to_bytes(flat), "dataset_key": "training"}, timeout=1e-3)
        with pytest.raises(ConnectionError):
            cl.ping()
        deadline = time.time() + 20                          # let the orphaned fit finish before tearing down
        while w.fits_served < 4 and time.time() < deadline:
            time.sleep(0.01)
        assert w.fits_served == 4
        # a frame header announcing more than MAX_FRAME bytes is refused and the connection dropped
        s = socket.create_connection(("127.0.0.1", p), timeout=5)
        s.sendall(struct.pack(">I", W.MAX_FRAME + 1))
        s.settimeout(5)
        assert s.recv(1) == b""
        s.close()
        cl2 = RemoteWorkerClient("y", "127.0.0.1", p)       # the server itself is still fine
        assert cl2.ping()
        cl2.close()
    finally:
        w.stop()


def test_duplicate_and_mid_training_announcements(tmp_path):
    """A duplicated TRAINING event must not leak the first connection; a device that announces itself again while
    its own training is running is a late joiner for the *next* window, not wiped by the end-of-training cleanup."""
    w1, p1 = _worker(seed=1)
    w2, p2 = _worker(seed=2)
    try:
        c, pub, clock = make(tmp_path, remote=True, rounds=1, args=Arguments(lr=0.05))
        ident = f"127.0.0.1:{p1}"
        pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")
        c.drain()
        first = c.known_workers[ident]
        pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")      # duplicate delivery
        pub.publish(TOPIC, f"(127.0.0.1, {p2}, TRAINING)")
        c.drain()
        second = c.known_workers[ident]
        assert second is not first and first._sock is None and second._sock is not None
        assert len(settings.training_devices) == 2

        # re-announce from inside the training (the fit RPC of worker 2 triggers it)
        orig_fit = RemoteWorkerClient.fit
        fired = []

        def fit_and_reannounce(self, *a, **k):
            if not fired and self.id == ident:
                fired.append(1)
                pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")
                c.drain()
            return orig_fit(self, *a, **k)

        RemoteWorkerClient.fit = fit_and_reannounce
        try:
            clock.advance(1.0)
        finally:
            RemoteWorkerClient.fit = orig_fit
        res = c.windower.last_result
        assert sorted(res["workers"]) == sorted([ident, f"127.0.0.1:{p2}"]) and res["dropped"] == []
        # worker 2 is forgotten, worker 1's newer registration survived and armed the next window
        assert list(settings.training_devices) == [ident] and ident in c.known_workers
        assert c.known_workers[ident] is not second and second._sock is None
        # reference semantics: the late joiner waits for the next event to arm a window (fc.py:187-191)
        assert c.windower.state == "IDLE" and c.trainings_done == 1
        pub.publish(TOPIC, f"(127.0.0.1, {p2}, TRAINING)")
        c.drain()
        assert c.windower.state == "COLLECTING"
        clock.advance(1.0)
        assert c.trainings_done == 2 and sorted(c.windower.last_result["workers"]) == sorted([ident, f"127.0.0.1:{p2}"])
        assert len(settings.training_devices) == 0
        c.shutdown()
    finally:
        w1.stop(); w2.stop()


def test_remote_round_fans_out_to_all_devices_at_once(tmp_path):
    """40 devices whose fit takes 0.25 s each: the round must take about one fit, not 40 / (cpu_count + 4) of them
    (the default asyncio executor would serialise the upper-bound-100 case the reference advertises)."""
    from collections import OrderedDict

    class SlowWorker:
        def __init__(self, wid):
            self.id, self.closed = wid, False

        def fit(self, flat, cfg, dataset_key="training", timeout=None):
            time.sleep(0.25)
            return flat.clone() + 0.01, 0.5, 10

        def close(self):
            self.closed = True

    c, pub, clock = make(tmp_path, remote=True, rounds=2)
    snapshot = OrderedDict((f"10.0.0.{i}:8777", SlowWorker(f"10.0.0.{i}:8777")) for i in range(1, 41))
    for wid, w in snapshot.items():
        c.registry.register(wid, w)
    t0 = time.time()
    res = c._start_training(snapshot)
    dt = time.time() - t0
    assert len(res["losses"]) == 40 and res["dropped"] == [] and all(w.closed for w in snapshot.values())
    assert dt < 1.5, dt      # 2 rounds x 0.25 s + overhead; behind the default executor (cpu_count + 4 threads): >= 2 s here
    assert len(settings.training_devices) == 0


def test_silent_remote_worker_is_dropped_after_the_fit_timeout(tmp_path):
    """SURVEY §4 fault injection: a device that accepts the fit request and never answers ("edge devices do not fail in
    the training phases" is an assumption of the paper, p.6; the reference would block forever in ``async_fit``).  With
    ``--fit-timeout`` the round finishes with the devices that did answer and the silent one is dropped for good."""
    import socket
    import threading

    good, pg = _worker(seed=1)
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(4)
    ps = srv.getsockname()[1]
    held = []

    def black_hole():
        while True:
            try:
                conn, _ = srv.accept()
            except OSError:
                return
            held.append(conn)                                  # read nothing, answer nothing, keep the socket open

    threading.Thread(target=black_hole, daemon=True).start()
    try:
        c, pub, clock = make(tmp_path, remote=True, rounds=2, fit_timeout=0.5)
        pub.publish(TOPIC, f"(127.0.0.1, {pg}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {ps}, TRAINING)")
        c.drain()
        assert len(settings.training_devices) == 2
        t0 = time.time()
        clock.advance(1.0)
        took = time.time() - t0
        res = c.windower.last_result
        assert res["dropped"] == [f"127.0.0.1:{ps}"] and list(res["losses"]) == [f"127.0.0.1:{pg}"]
        assert good.fits_served == 2                           # the second round ran with the surviving device only
        assert 0.4 < took < 5.0                                # one timeout, not one per round
        assert os.path.exists(c.path) and len(settings.training_devices) == 0
    finally:
        good.stop()
        srv.close()
        for s in held:
            s.close()


def test_prometheus_exporter_serves_round_metrics(tmp_path):
    """``--metrics-port``: rounds, round time, per-device loss and traffic as Prometheus series."""
    import urllib.request

    from colearn_federated_learning_b200.utils.metrics import PrometheusExporter, RoundLogger

    w1, p1 = _worker(seed=1)
    w2, p2 = _worker(seed=2)
    exp = PrometheusExporter(0)
    try:
        c, pub, clock = make(tmp_path, remote=True, rounds=3, metrics=RoundLogger(exporter=exp))
        pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {p2}, TRAINING)")
        c.drain()
        clock.advance(1.0)
        body = urllib.request.urlopen(f"http://127.0.0.1:{exp.port}/metrics", timeout=5).read().decode()
        series = dict(line.rsplit(" ", 1) for line in body.splitlines() if line and not line.startswith("#"))
        assert float(series["colearn_rounds_total"]) == 3 and float(series["colearn_trainings_total"]) == 1
        assert float(series["colearn_round_seconds_count"]) == 3 and float(series["colearn_round_selected_workers"]) == 2
        assert float(series["colearn_bytes_to_workers_total"]) > 3 * 2 * 9604
        assert float(series["colearn_bytes_from_workers_total"]) > 3 * 2 * 9604
        assert f'colearn_worker_last_loss{{worker="127.0.0.1:{p1}"}}' in series
    finally:
        exp.close()
        w1.stop(); w2.stop()


def test_small_model_sections_run_with_one_intra_op_thread():
    """utils/threads.py: the coordinator's training section for a few-thousand-parameter model must not pay OpenMP
    fork/join per op; the previous thread setting is restored afterwards, large models keep the pool."""
    from colearn_federated_learning_b200.utils.threads import small_model_threads
    before = torch.get_num_threads()
    if before == 1:
        pytest.skip("single-threaded torch build / environment")
    with small_model_threads(4866, torch.device("cpu")):
        assert torch.get_num_threads() == 1
    assert torch.get_num_threads() == before
    with small_model_threads(11_000_000, torch.device("cpu")):
        assert torch.get_num_threads() == before
    with small_model_threads(4866, torch.device("cuda", 0)):    # host-side ops of a GPU-resident small model are small too
        assert torch.get_num_threads() == 1
    try:
        with small_model_threads(10, None):
            raise RuntimeError("x")
    except RuntimeError:
        pass
    assert torch.get_num_threads() == before


def test_remote_mode_returns_and_averages_batchnorm_statistics(tmp_path):
    """The flat arena carries parameters only; a device's BatchNorm running statistics come back next to them and the
    saved global model holds their FedAvg mean instead of the initial (0, 1) statistics."""
    from torch import nn
    from colearn_federated_learning_b200.models import register_model

    class BnToy(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.bn, self.fc2 = nn.Linear(10, 8), nn.BatchNorm1d(8), nn.Linear(8, 2)

        def forward(self, x):
            return self.fc2(torch.relu(self.bn(self.fc1(x))))

    register_model("bn_toy", BnToy, default_loss="xent", overwrite=True)
    w1, p1 = _worker(seed=1)
    w2, p2 = _worker(seed=2)
    try:
        a = Arguments(lr=0.05)
        a.model, a.batch_size = "bn_toy", 16
        c, pub, clock = make(tmp_path, remote=True, rounds=2, args=a)
        pub.publish(TOPIC, f"(127.0.0.1, {p1}, TRAINING)")
        pub.publish(TOPIC, f"(127.0.0.1, {p2}, TRAINING)")
        c.drain()
        clock.advance(1.0)
        res = c.windower.last_result
        assert res["dropped"] == [] and len(res["losses"]) == 2
        state = torch.load(c.path, weights_only=True)
        assert float(state["bn.running_mean"].abs().max()) > 0 and not torch.allclose(state["bn.running_var"], torch.ones(8))
        devs = [w._models["bn_toy"][0].state_dict() for w in (w1, w2)]
        want = (devs[0]["bn.running_mean"] + devs[1]["bn.running_mean"]) / 2          # equal shards -> equal weights
        assert torch.allclose(state["bn.running_mean"], want.cpu(), atol=1e-6)
    finally:
        w1.stop(); w2.stop()
