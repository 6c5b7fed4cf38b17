#!/usr/bin/env python
"""CoLearn device-side worker CLI (flag-compatible with the reference ``remote_worker.py:18-47``).

    python remote_worker.py --host 127.0.0.1 -p 8778 -b localhost -t "topic/state" -w 1 \
        -e "TRAINING" --verbose -dt <training csv> -di <inference csv>

The host MUST be given as an IP (it is what the coordinator parses out of the event).  The worker
hosts its dataset under key ``"training"`` (rw.py:108), optional inference tensors tagged
``"inference"`` (rw.py:102-104), announces itself on the bus after ``--wait`` seconds
(rw.py:110-114) and then serves fit/search/predict RPCs forever (rw.py:117).
"""
import argparse
import logging
import os
import sys
from threading import Timer

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Run a federated worker (RPC server).")
    parser.add_argument("--port", "-p", type=int, default=8777, help="port number of the worker server, e.g. --port 8777")
    parser.add_argument("--host", type=str, required=True, help="ip address of the interface the worker listens on")
    parser.add_argument("--broker", "-b", type=str, required=True, help="bus broker host")
    parser.add_argument("--topic", "-t", type=str, required=True, help="topic where the event must be published")
    parser.add_argument("--wait", "-w", type=int, default=5, help="seconds to wait before sending the event")
    parser.add_argument("--event", "-e", type=str, default="TRAINING", help="state of the client (TRAINING, INFERENCE, NOT_READY)")
    parser.add_argument("--training", "-dt", type=str, default=None, help="training data csv")
    parser.add_argument("--inference", "-di", type=str, default=None, help="inference data csv")
    parser.add_argument("--verbose", "-v", action="store_true", help="verbose worker")
    # extensions
    parser.add_argument("--broker-port", type=int, default=1883)
    parser.add_argument("--synthetic", type=int, default=0, help="host N synthetic rows instead of --training")
    parser.add_argument("--data-model", default="ffnn",
                        help="architecture the synthetic data is shaped for (resnet18: 32x32 images, net: 28x28, "
                             "testing_remote: 2 features, otherwise the ten UNSW-IoT features)")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--no-cuda", action="store_true")
    return parser


def main(args: argparse.Namespace) -> None:  # pragma: no cover - exercised by the CLI integration test
    import torch

    from colearn_federated_learning_b200.control.bus import BusClient
    from colearn_federated_learning_b200.control.event_parser import format_event
    from colearn_federated_learning_b200.control.workers import WorkerServer
    from colearn_federated_learning_b200.data import (BaseDataset, NetworkTrafficDataset, synthetic_for_model,
                                                      xor_toy_dataset)

    logging.basicConfig(format="%(asctime)s: %(message)s", level=logging.INFO, datefmt="%H:%M:%S")
    identifier = args.host + ":" + str(args.port)
    device = torch.device("cpu" if args.no_cuda or not torch.cuda.is_available() else "cuda")

    # unique client id per worker (the reference's literal "woker" gets duplicates kicked)
    client = BusClient(client_id="worker-" + identifier, transport="tcp")
    # last will: if this process dies without a DISCONNECT the broker withdraws the device for us
    client.will_set(args.topic, format_event(args.host, "NOT_READY", args.port))
    client.connect(args.broker, args.broker_port)
    to_publish = format_event(args.host, args.event, args.port)

    if args.synthetic > 0:
        dataset = BaseDataset(*synthetic_for_model(args.data_model, args.synthetic, seed=args.seed))
    elif args.training is None:
        dataset = xor_toy_dataset()  # rw.py:75-80
    else:
        print(args.training)
        dataset = NetworkTrafficDataset(args.training)

    worker = WorkerServer(identifier, args.host, args.port, device=device, verbose=args.verbose)
    if args.inference is not None:
        print(args.inference)
        dataset_inf = NetworkTrafficDataset(args.inference)
        worker.load_data([torch.tensor(row).float() for row in dataset_inf.data], tag="inference")
    worker.add_dataset(dataset, key="training")

    t = Timer(args.wait, lambda: client.publish(args.topic, to_publish))
    t.daemon = True
    t.start()
    try:
        worker.start()  # blocks forever
    except KeyboardInterrupt:
        worker.stop()


if __name__ == "__main__":
    main(build_parser().parse_args())
