"""Evaluate a saved global model — what the reference's unused ``evaluate()`` helper was for
(``client_federated.py:217-253``; its only call site is commented out at ``federated_coordinator.py:587-590``).

    python -m colearn_federated_learning_b200.tools.evaluate_model --checkpoint test.pth \
        [--model ffnn] [--test-path data.csv | --synthetic 1000] [--json]

The architecture is taken from the checkpoint's JSON sidecar when there is one, else from ``--model``; a checkpoint
that does not fit the architecture is an error (exit code 2), not a silent random-init evaluation.
"""
import argparse
import json
import sys

import torch

from ..data import NetworkTrafficDataset, synthetic_for_model
from ..fl.evaluate import evaluate
from ..models import MODEL_REGISTRY, build_model
from ..utils.checkpoint import checkpoint_compatible, load_meta, load_or_init


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--checkpoint", default="./test.pth")
    ap.add_argument("--model", choices=sorted(MODEL_REGISTRY), default=None)
    ap.add_argument("--test-path", default=None, help="Bot-IoT CSV (the reference's evaluation data)")
    ap.add_argument("--synthetic", type=int, default=0, help="evaluate on N synthetic rows shaped for the model")
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--json", action="store_true", help="print one JSON object instead of the reference's text line")
    ns = ap.parse_args(argv)

    name = ns.model or load_meta(ns.checkpoint).get("model") or "ffnn"
    model = build_model(name)
    if not checkpoint_compatible(model, ns.checkpoint):
        print(f"{ns.checkpoint} is missing or does not hold a {name!r} model", file=sys.stderr)
        return 2
    load_or_init(model, ns.checkpoint)
    device = torch.device("cuda" if torch.cuda.is_available() and not ns.no_cuda else "cpu")
    if ns.synthetic > 0 or ns.test_path is None:
        x, y = synthetic_for_model(name, ns.synthetic or 1000, seed=ns.seed)
    else:
        x, y = NetworkTrafficDataset(ns.test_path).tensors()
    res = evaluate(model.to(device), x.to(device), y.to(device), verbose=not ns.json)
    if ns.json:
        print(json.dumps({"model": name, "checkpoint": ns.checkpoint, **res}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
