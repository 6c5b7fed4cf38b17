"""Selection bounds, FedAvg, checkpoint, data layer, models, tooling (all CPU)."""
import os
import subprocess

import numpy as np
import pytest
import torch
from hypothesis import given, settings as hsettings, strategies as st

from colearn_federated_learning_b200 import ops, settings
from colearn_federated_learning_b200.control.arguments import Arguments
from colearn_federated_learning_b200.control.selection import SelectionPolicy, encrypted_policy
from colearn_federated_learning_b200.data import (FEATURE_COLUMNS, CSV_HEADER, FederatedDataLoader, NetworkTrafficDataset,
                                                  Normalize, ToTensor, ToTensorLong, federate, minmax_scale,
                                                  shard_bounds, synthetic_unsw, write_synthetic_csv, xor_toy_dataset)
from colearn_federated_learning_b200.fl import federated_avg, federated_avg_flat, normalized_weights
from colearn_federated_learning_b200.models import (FFNN, MLP, Net, ResNet18, TestingRemote, WIDE_MLP_SPEC,
                                                    alias_params_to_arena, build_model, flatten_params, num_params,
                                                    state_dict_from_flat, unflatten_params)
from colearn_federated_learning_b200.utils.checkpoint import (checkpoint_compatible, load_meta, load_or_init, save_model)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- selection ---------------------------------------------------------------------------------
def test_selection_bounds_and_policies():
    devs = {f"d{i}": i for i in range(8)}
    assert list(SelectionPolicy().select(devs)) == list(devs)                       # reference: all
    assert SelectionPolicy(lower_bound=9).select(devs) == {}
    assert list(SelectionPolicy(select_k=4, policy="first").select(devs)) == ["d0", "d1", "d2", "d3"]
    r1 = list(SelectionPolicy(select_k=4, policy="random", seed=3).select(devs, 0))
    assert len(r1) == 4 and r1 == list(SelectionPolicy(select_k=4, policy="random", seed=3).select(devs, 0))
    assert list(encrypted_policy().select(devs)) == ["d0", "d1"]                    # fc.py:401-404
    assert encrypted_policy().select({"only": 1}) == {}                             # needs >= 2
    big = {f"d{i}": i for i in range(120)}
    assert len(SelectionPolicy().select(big)) == 120                                # reference only logs at upper bound
    assert len(SelectionPolicy(policy="first").select(big)) == 100


def test_arguments_defaults_match_reference():
    a = Arguments()
    assert (a.batch_size, a.test_batch_size, a.epochs, a.federate_after_n_batches, a.lr, a.momentum, a.seed,
            a.log_interval) == (1, 1024, 1, -1, 0.01, 0.5, 1, 30)
    a.set_federated_batches(1000)
    assert a.federate_after_n_batches == 1000


def test_settings_compat_surface():
    reg = settings.init()
    settings.training_devices["a"] = 1
    assert "a" in settings.training_devices and settings.event_served == 0
    reg.serve_event()
    assert settings.event_served == 1
    copy = settings.training_devices.copy()
    del settings.training_devices["a"]
    assert copy == {"a": 1} and len(settings.training_devices) == 0


# ---- FedAvg --------------------------------------------------------------------------------------
def test_federated_avg_uniform_matches_naive_mean_and_is_in_place():
    torch.manual_seed(0)
    models = {f"w{i}": FFNN() for i in range(3)}
    want = torch.stack([flatten_params(m) for m in models.values()]).mean(0)
    out = federated_avg(models)
    assert out is models["w0"]
    assert torch.allclose(flatten_params(out), want, atol=1e-7)


def test_federated_avg_weighted_and_alias_guard():
    torch.manual_seed(1)
    a, b = MLP(), MLP()
    fa, fb = flatten_params(a).clone(), flatten_params(b).clone()
    federated_avg({"a": a, "b": b}, sample_counts={"a": 300, "b": 100})
    assert torch.allclose(flatten_params(a), 0.75 * fa + 0.25 * fb, atol=1e-7)
    m = MLP()
    with pytest.raises(ValueError):
        federated_avg({"x": m, "y": m})     # the reference's aliasing bug is rejected


def test_fedavg_buffers_not_averaged_by_default():
    torch.manual_seed(2)
    a, b = ResNet18(), ResNet18()
    a.bn1.running_mean.fill_(1.0)
    b.bn1.running_mean.fill_(3.0)
    federated_avg({"a": a, "b": b})
    assert torch.allclose(a.bn1.running_mean, torch.ones(64))
    a.bn1.running_mean.fill_(1.0)
    federated_avg({"a": a, "b": b}, average_buffers=True)
    assert torch.allclose(a.bn1.running_mean, torch.full((64,), 2.0))


@given(st.integers(2, 6), st.integers(0, 10 ** 6))
@hsettings(max_examples=20, deadline=None)
def test_fedavg_flat_linearity_and_permutation_invariance(k, seed):
    g = torch.Generator().manual_seed(seed)
    flats = [torch.randn(50, generator=g) for _ in range(k)]
    counts = [int(c) for c in torch.randint(1, 100, (k,), generator=g)]
    base = federated_avg_flat(flats, counts)
    perm = torch.randperm(k, generator=g).tolist()
    assert torch.allclose(base, federated_avg_flat([flats[i] for i in perm], [counts[i] for i in perm]), atol=1e-5)
    assert torch.allclose(federated_avg_flat([2 * f for f in flats], counts), 2 * base, atol=1e-5)
    assert torch.allclose(federated_avg_flat(flats, [1] * k), federated_avg_flat(flats, None), atol=1e-6)
    assert abs(float(normalized_weights(counts, k).sum()) - 1.0) < 1e-6


# ---- checkpoint --------------------------------------------------------------------------------------
def test_checkpoint_roundtrip_reference_keys(tmp_ckpt):
    m = FFNN()
    save_model(m, tmp_ckpt, meta={"rounds": 3})
    state = torch.load(tmp_ckpt, weights_only=True)
    assert list(state) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias",
                           "fc4.weight", "fc4.bias"]
    m2 = FFNN()
    assert load_or_init(m2, tmp_ckpt) is True
    assert torch.equal(flatten_params(m), flatten_params(m2))
    assert load_meta(tmp_ckpt)["rounds"] == 3
    assert load_or_init(FFNN(), tmp_ckpt + ".missing") is False
    assert checkpoint_compatible(FFNN(), tmp_ckpt) and not checkpoint_compatible(TestingRemote(), tmp_ckpt)


def test_flat_arena_roundtrip_and_alias():
    m = MLP()
    flat = flatten_params(m).clone()
    m2 = unflatten_params(MLP(), flat)
    assert torch.equal(flatten_params(m2), flat)
    arena = torch.zeros(num_params(m))
    alias_params_to_arena(m, arena)
    assert torch.equal(arena, flat)
    arena.add_(1.0)                                   # kernels updating the arena update the module
    assert torch.equal(flatten_params(m), flat + 1.0)
    sd = state_dict_from_flat(m, arena)
    assert torch.equal(sd["fc1.weight"], m.fc1.weight.detach())


def test_model_zoo_param_counts():
    assert [num_params(build_model(n)) for n in ("ffnn", "testing_remote", "net", "mlp", "resnet18")] == \
        [2401, 671, 109386, 4994, 11181642]
    assert WIDE_MLP_SPEC.n_params == 50397186
    assert FFNN()(torch.zeros(3, 10)).shape == (3, 1) and Net()(torch.zeros(2, 1, 28, 28)).shape == (2, 10)
    assert ResNet18()(torch.zeros(2, 3, 32, 32)).shape == (2, 10)
    assert FFNN().get_traced_model()(torch.zeros(10)).shape == (1,)


# ---- data ----------------------------------------------------------------------------------------------
def test_network_traffic_dataset(tmp_path):
    p = str(tmp_path / "d.csv")
    write_synthetic_csv(p, 37, seed=3)
    assert open(p).readline().strip().split(",") == CSV_HEADER
    ds = NetworkTrafficDataset(p, transform=ToTensor(torch.device("cpu")))
    assert len(ds) == 37
    x, y = ds[5]
    assert x.shape == (10,) and x.dtype == torch.float32 and y.shape == (1,)
    X, Y = ds.tensors()
    assert X.shape == (37, 10) and float(X.min()) == 0.0 and float(X.max()) == 1.0 and Y.shape == (37, 1)
    import pandas as pd
    df = pd.read_csv(p)
    assert np.allclose(X.numpy(), minmax_scale(df[FEATURE_COLUMNS].values), atol=1e-6)   # column order = ds.py:29


def test_transforms():
    t = Normalize()(torch.tensor([1.0, 3.0]))
    assert torch.allclose(t, torch.tensor([-0.5, 0.5]))
    assert ToTensorLong(torch.device("cpu"))(np.array([1.7])).dtype == torch.int64
    assert minmax_scale(np.array([[1.0, 5.0], [1.0, 7.0]]))[:, 0].tolist() == [0.0, 0.0]   # constant column -> 0


def test_federate_contract():
    assert shard_bounds(10, 3) == [(0, 4), (4, 8), (8, 10)]            # ceil(N/K) contiguous, in order
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    x, y = synthetic_unsw(10, seed=0)
    from colearn_federated_learning_b200.data import BaseDataset
    fed = federate(BaseDataset(x, y), ["a", "b", "c"])
    assert [len(fed[w]) for w in fed.workers] == [4, 4, 2] and torch.equal(fed["b"].x, x[4:8])
    seen = {}
    for wid, data, target in FederatedDataLoader(fed, batch_size=3, shuffle=True):
        seen.setdefault(wid, []).append(data)
    assert list(seen) == ["a", "b", "c"]                                # worker by worker
    assert sorted(torch.cat(seen["a"]).sum(1).tolist()) == sorted(x[0:4].sum(1).tolist())
    assert len(xor_toy_dataset()) == 4


# ---- CPU ops dispatch -----------------------------------------------------------------------------------------
def test_ops_cpu_dispatch():
    z, y = torch.randn(20, 1), (torch.rand(20, 1) > 0.5).float()
    l, dz = ops.sigmoid_bce(z, y)
    assert torch.allclose(l, torch.nn.functional.binary_cross_entropy(torch.sigmoid(z), y), atol=1e-6)
    logits, labels = torch.randn(9, 4), torch.randint(0, 4, (9,))
    l, d = ops.softmax_xent(logits, labels)
    assert torch.allclose(l, torch.nn.functional.cross_entropy(logits, labels), atol=1e-6)
    assert torch.equal(ops.argmax_rows(logits), logits.argmax(1, keepdim=True))
    p = torch.randn(10)
    ops.sgd_step(p, torch.ones(10), 0.5)
    assert ops.device_permutation(11, 2, 0, "cpu").shape == (2, 11)
    s = ops.minmax_scale(torch.rand(5, 3))
    assert float(s.min()) == 0.0 and float(s.max()) == 1.0


# ---- tooling ------------------------------------------------------------------------------------------------------
def test_file_upgrader_new_and_del(tmp_path):
    from colearn_federated_learning_b200.tools import file_upgrader as fu
    f = str(tmp_path / "filtering_file.txt")
    assert fu.main(["-c", "NEW", "-i", "10.0.0.1", "-f", f]) == 0
    fu.main(["-c", "NEW", "-i", "10.0.0.2", "-f", f])
    fu.main(["-c", "NEW", "-i", "10.0.0.1", "-f", f])               # duplicate ignored
    assert fu.read_ips(f) == ["10.0.0.1", "10.0.0.2"]
    fu.main(["-c", "DEL", "-i", "10.0.0.1", "-f", f])               # really deletes (reference bug 2.8-10)
    assert fu.read_ips(f) == ["10.0.0.2"]
    assert fu.main(["-c", "WAT", "-i", "1.1.1.1", "-f", f]) == 1


def test_monitoring_dnsmasq_dry_run(tmp_path):
    logf = tmp_path / "dhcpmasq.txt"
    logf.write_text("t|NEW|a|b|c|d|https://mud.example/x.json|e|f|192.168.1.44\n"
                    "t|OLD|a|b|c|d|https://mud.example/x.json|e|f|192.168.1.45\n"
                    "t|DEL|a|b|c|d|-|e|f|192.168.1.46\n"
                    "t|DEL|a|b|c|d|https://mud.example/y.json|e|f|192.168.1.47\n")
    out = subprocess.run(["sh", os.path.join(ROOT, "device_filtering", "monitoring_dnsmasq.sh"), "-u", "me", "-p", "/opt/co",
                          "-l", str(logf), "-n", "-1"], capture_output=True, text=True, timeout=30).stdout
    assert "file_upgrader.py -c NEW -i 192.168.1.44" in out and "-c DEL -i 192.168.1.47" in out
    assert "192.168.1.45" not in out.replace("not valid: ", "#").split("ssh")[-1] and out.count("ssh ") == 2


def test_feature_generator(tmp_path):
    import pandas as pd
    from colearn_federated_learning_b200.tools.feature_generator import MODEL_FEATURES, generate
    rows = []
    rng = np.random.default_rng(0)
    for i in range(30):
        rows.append(dict(SrcAddr=f"10.0.0.{i % 3}", DstAddr=f"10.0.1.{i % 2}", Sport=1000 + i if i % 5 else np.nan,
                         Dport=80, Proto="tcp" if i % 2 else "udp", State=["REQ", "CON", "FIN", "XYZ"][i % 4],
                         SrcBytes=100 + i, DstBytes=50 + i, SrcPkts=3, DstPkts=2, TotPkts=5, TotBytes=150 + 2 * i,
                         SrcRate=1.5, DstRate=0.5, Dur=0.5 + 0.1 * i, Seq=i, StdDev=0.1, Min=0.0, Max=1.0, Mean=0.5))
    src, dst = str(tmp_path / "in.csv"), str(tmp_path / "out.csv")
    pd.DataFrame(rows).to_csv(src, index=False)
    df = generate(src, dst, extract=False)
    assert set(MODEL_FEATURES) <= set(df.columns) and df["state_number"].isin([1, 2, 3, 4, 5, 6, -1]).all()
    assert (df.loc[df["state"] == "XYZ", "state_number"] == -1).all() and (df["sport"].notna()).all()
    n_in = df[df["saddr"] == "10.0.0.0"]["N_IN_Conn_P_SrcIP"].iloc[0]
    assert n_in == sum(1 for r in rows if r["SrcAddr"] == "10.0.0.0" and r["State"] in ("REQ", "CON", "EST"))
    out = generate(src, dst, extract=True)
    assert list(out.columns) == MODEL_FEATURES and os.path.exists(dst)


def test_monitors(tmp_path):
    from colearn_federated_learning_b200.utils import monitors
    cpu, net, temp = (str(tmp_path / n) for n in ("cpu.txt", "net.txt", "temp.txt"))
    monitors.monitor_cpu(os.getpid(), cpu, samples=2, interval=0.01)
    assert open(cpu).read().count("Monitoring: ") == 2
    import psutil
    iface = next(iter(psutil.net_io_counters(pernic=True)))
    monitors.monitor_network(iface, net, samples=2, interval=0.01)
    txt = open(net).read()
    assert "incoming       : bytes=0B, pkts=0" in txt and txt.count("Monitoring: ") == 2
    monitors.monitor_temperature(temp, interval=0.01, samples=2)
    assert open(temp).read().count("Monitoring: ") == 2
    s = monitors.NvmlSampler().start()
    assert "reasons" in s.stop()
    gpu = str(tmp_path / "gpu.txt")
    monitors.monitor_gpu(0, gpu, samples=1, interval=0.01)      # no NVML on the CPU box -> says so, never raises
    assert os.path.getsize(gpu) > 0
    ns = monitors.build_parser().parse_args(["-p", "1", "-n", "eth0", "--gpu", "0"])
    assert (ns.pid, ns.network, ns.gpu) == (1, "eth0", 0)
    # NVLink payload counters: without NVML / NVLink the object says so and its helpers stay total (bench.py / comm_sweep.py branch on .ok)
    c = monitors.NvlinkCounters(index=0, uuid="GPU-00000000-0000-0000-0000-000000000000")
    if not c.ok:
        assert c.err and monitors.NvlinkCounters.delta(None, {"tx_bytes": 1, "rx_bytes": 2}) is None
    assert monitors.NvlinkCounters.delta({"tx_bytes": 1024, "rx_bytes": 0}, {"tx_bytes": 4096, "rx_bytes": 2048}) == {"tx_bytes": 3072, "rx_bytes": 2048}


def test_round_overhead_fit_separates_slope_and_intercept():
    """scripts/round_overhead.py: per-step time and fixed cost of a round are the slope and intercept over steps per round."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("round_overhead", os.path.join(ROOT, "scripts", "round_overhead.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    slope, icpt = mod.fit_line([(n, 0.000562 * n + 0.0245) for n in (128, 512, 1024)])
    assert abs(slope - 0.000562) < 1e-9 and abs(icpt - 0.0245) < 1e-9


def test_reference_shaped_client_federated_api(tmp_path):
    from colearn_federated_learning_b200 import client_federated as cf
    from colearn_federated_learning_b200.data import BaseDataset
    pred, tgt = torch.tensor([[0.8], [0.3]]), torch.tensor([[1.0], [0.0]])
    assert torch.allclose(cf.loss_fn(target=tgt, pred=pred), torch.nn.functional.binary_cross_entropy(pred, tgt))
    x, y = synthetic_unsw(40, seed=0)
    fed = federate(BaseDataset(x, y), ["a", "b"])
    loader = FederatedDataLoader(fed, batch_size=1, shuffle=True)
    a = Arguments()
    m = cf.FFNN()
    before = flatten_params(m).clone()
    m2, loss = cf.train_local("a", m, torch.optim.SGD(m.parameters(), lr=0.05), 1, loader, a)
    assert m2 is m and not torch.equal(flatten_params(m), before) and loss.ndim == 0
    res = cf.evaluate(cf.FFNN(), [(x[i:i + 8], y[i:i + 8]) for i in range(0, 40, 8)], torch.device("cpu"))
    assert res["n"] == 40 and 0 <= res["accuracy"] <= 1
    shared = cf.get_private_data_loaders(["a", "b"], a, n_train_items=5, dataset=BaseDataset(x, y))
    assert len(shared) == 5


@pytest.mark.parametrize("name", ["ffnn", "testing_remote", "net", "mlp", "resnet18"])
def test_synthetic_data_matches_every_architecture(name):
    """``--synthetic N`` must produce data the chosen ``--model`` can train on (box mode / local mode / workers)."""
    from colearn_federated_learning_b200.data import synthetic_for_model
    from colearn_federated_learning_b200.fl.trainer import FitConfig, local_fit
    from colearn_federated_learning_b200.models import build_model, flatten_params

    x, y = synthetic_for_model(name, 24, seed=2)
    assert x.shape[0] == 24 and y.shape == (24, 1)
    model = build_model(name)
    flat = flatten_params(model).clone()
    before = flat.clone()
    loss, _ = local_fit(flat, model, x, y, FitConfig(model=name, loss="auto", batch_size=8, lr=0.05))
    assert torch.isfinite(loss) and torch.isfinite(flat).all() and not torch.equal(flat, before)


def test_osmud_probe_dry_run(tmp_path):
    """``data/osmud_test.sh`` (router-side osMUD probe): named options, the reference's positional form, dry-run."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["sh", os.path.join(root, "data", "osmud_test.sh"), "-d", "-n", "2", "-t", "3", "-o", str(tmp_path / "r")],
                         capture_output=True, text=True, check=True).stdout
    assert out.count("/etc/init.d/firewall restart") == 2 and "+ sleep 3" in out and f"{tmp_path}/r/test_2.txt" in out
    assert not (tmp_path / "r").exists()
    out = subprocess.run(["sh", os.path.join(root, "data", "osmud_test.sh"), "-d", "1", "7"], capture_output=True, text=True,
                         check=True, cwd=str(tmp_path)).stdout
    assert "+ sleep 7" in out and "run 1/1" in out


def test_evaluate_model_tool(tmp_path, capsys):
    """``tools/evaluate_model``: the reference's (unused) evaluate() as a CLI over a saved checkpoint."""
    import json

    from colearn_federated_learning_b200.models import MLP, FFNN
    from colearn_federated_learning_b200.tools import evaluate_model
    from colearn_federated_learning_b200.utils.checkpoint import save_model

    ck = str(tmp_path / "m.pth")
    save_model(MLP(), ck, meta={"model": "mlp", "rounds": 1})
    assert evaluate_model.main(["--checkpoint", ck, "--synthetic", "64", "--json", "--no-cuda"]) == 0
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert res["model"] == "mlp" and res["n"] == 64 and 0 <= res["accuracy"] <= 1 and res["loss"] > 0
    # text mode prints the reference's line; a mismatching architecture is refused
    save_model(FFNN(), ck)
    assert evaluate_model.main(["--checkpoint", ck, "--model", "ffnn", "--synthetic", "32", "--no-cuda"]) == 0
    assert "Test set: Average loss:" in capsys.readouterr().out
    assert evaluate_model.main(["--checkpoint", ck, "--model", "mlp", "--no-cuda"]) == 2
    assert evaluate_model.main(["--checkpoint", str(tmp_path / "missing.pth"), "--no-cuda"]) == 2
