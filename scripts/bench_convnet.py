"""ResNet-18 local fit: this repo's conv path (eager / CUDA graph) vs the autograd + cuDNN path, one GPU.

    python scripts/bench_convnet.py [--samples 2048] [--batch 128] [--reps 3]

Device-timed with CUDA events around whole fits (16 SGD steps at the defaults); prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.fl.convnet import ConvNetTrainer  # noqa: E402
from colearn_federated_learning_b200.fl.trainer import FitConfig, local_fit  # noqa: E402
from colearn_federated_learning_b200.models.registry import flatten_params  # noqa: E402
from colearn_federated_learning_b200.models.resnet import ResNet18  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="", help="comma-separated subset of torch_cudnn_autocast,native_eager,native_graph")
    args = ap.parse_args()
    only = {t for t in args.only.split(",") if t}
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(args.samples, 3, 32, 32, device=dev)
    y = torch.randint(0, 10, (args.samples, 1), device=dev).float()
    cfg = FitConfig(model="resnet18", loss="xent", batch_size=args.batch, epochs=1, lr=0.01, shuffle=True, seed=1)
    steps = args.samples // args.batch
    out = {"steps_per_fit": steps, "batch": args.batch,
           "flags": {k: v for k, v in os.environ.items() if k.startswith("COLEARN_CONV_") or k == "COLEARN_PDL"}}

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    for name, env, graph in (("torch_cudnn_autocast", "torch", None), ("native_eager", "native", False), ("native_graph", "native", True)):
        if only and name not in only:
            continue
        os.environ["COLEARN_CONV_PATH"] = env
        model = ResNet18(10).to(dev)
        flat = flatten_params(model)
        if env == "torch":
            fn = lambda: local_fit(flat, model, x, y, cfg)  # noqa: E731
        else:
            tr = ConvNetTrainer.cached(model, flat, args.batch, (32, 32))
            fn = lambda: tr.fit(flat, model, x, y.view(-1), cfg, None, use_graph=graph)  # noqa: E731
        ms = timed(fn)
        out[name] = {"ms_per_fit": ms, "ms_per_step": ms / steps}
        if env == "native":
            before = tr.launches
            fn()
            out[name]["launches_per_fit"] = tr.launches - before
        loss, _ = (fn() if env == "torch" else (fn(), None))
        out[name]["last_loss"] = float(loss)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
