"""Localise a schedule-dependent difference of the ResNet-18 step: per-parameter cosine / relative error of the update after
1 eager and 2 graph steps between the default schedule and the switches given on the command line.

    python scripts/debug_splitk_step.py COLEARN_CONV_SPLITK=1 [COLEARN_CONV_WGRAD_MN=1 ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.fl.convnet import ConvNetTrainer  # noqa: E402
from colearn_federated_learning_b200.models import flatten_params  # noqa: E402
from colearn_federated_learning_b200.models.resnet import ResNet18  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(256, 3, 32, 32, device=dev)
y = torch.randint(0, 10, (256,), device=dev)
FLAGS = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
KEYS = ["COLEARN_CONV_WGRAD_MN", "COLEARN_CONV_DGRAD_KN", "COLEARN_CONV_SPLITK", "COLEARN_CONV_STREAMS", "COLEARN_CONV_FUSED_BN",
        "COLEARN_CONV_SHADOW_T", "COLEARN_CONV_IMPLICIT"]


def run(steps, graph, perturb=0.0):
    """``perturb``: relative gaussian noise put on the parameters after step 1 (control: how far does a rounding-sized
    difference after one step move the update of two steps under the SAME schedule?)"""
    torch.manual_seed(1)
    model = ResNet18(10).to(dev)
    flat = flatten_params(model)
    flat0 = flat.clone()
    tr = ConvNetTrainer(model, dev, 128, (32, 32))
    tr.load(flat, model)
    for s in range(steps):
        lo = 128 * (s % 2)
        (tr._graph_step if graph else tr.step)(x[lo:lo + 128], y[lo:lo + 128], 0.05)
        if perturb and s == 0:
            tr.store(flat, model)
            g = torch.Generator(device=dev).manual_seed(7)
            flat.mul_(1.0 + perturb * torch.randn(flat.shape, device=dev, generator=g))
            tr.load(flat, model)
    tr.store(flat, model)
    torch.cuda.synchronize()
    return (flat - flat0).cpu(), model


out = {}
for k in KEYS:
    os.environ[k] = "0"
ctrl = {}
b2, _ = run(2, False)
for eps in (1e-6, 1e-5, 1e-4):
    p2, _ = run(2, False, eps)
    ctrl[str(eps)] = float((b2 * p2).sum() / (b2.norm() * p2.norm()))
out["control_same_schedule_perturbed_after_step1_cos"] = ctrl
for steps, graph in ((1, False), (2, False), (2, True)):
    for k in KEYS:
        os.environ[k] = "0"
    base, model = run(steps, graph)
    base2, _ = run(steps, graph)
    os.environ.update(FLAGS)
    got, _ = run(steps, graph)
    rows, off = [], 0
    for name, p in model.named_parameters():
        n = p.numel()
        b, g, b2 = base[off:off + n], got[off:off + n], base2[off:off + n]
        cos = float((b * g).sum() / (b.norm() * g.norm() + 1e-30))
        rel = float((b - g).norm() / (b.norm() + 1e-30))
        self_rel = float((b - b2).norm() / (b.norm() + 1e-30))
        rows.append({"name": name, "shape": list(p.shape), "cos": round(cos, 5), "rel_err": round(rel, 5), "rerun_rel_err": round(self_rel, 6)})
        off += n
    tot = float((base * got).sum() / (base.norm() * got.norm()))
    out[f"steps{steps}_{'graph' if graph else 'eager'}"] = {"total_cos": tot, "rerun_cos": float((base * base2).sum() / (base.norm() * base2.norm())),
                                                            "worst": sorted(rows, key=lambda r: r["cos"])[:4]}
print(json.dumps({"flags": FLAGS, **out}, indent=1))
