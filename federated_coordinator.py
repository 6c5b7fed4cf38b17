#!/usr/bin/env python
"""CoLearn coordinator CLI (flag-compatible with the reference ``federated_coordinator.py:64-90``).

    python federated_coordinator.py -t "topic/state"          # local (VirtualWorker) case
    python federated_coordinator.py -t "topic/state" -r       # remote case

Events are published to the bus as ``"(192.168.1.7, TRAINING)"`` (local) or
``"(127.0.0.1, 8777, TRAINING)"`` (remote) — e.g. with ``python -m
colearn_federated_learning_b200.tools.bus_pub -t topic/state -m "(192.168.1.7, TRAINING)"``, the
stand-in for ``mosquitto_pub``.  ``--host embedded`` (default when ``--host localhost`` has no
broker listening) starts the TCP bus broker inside this process.

On a multi-GPU box launch it under torchrun with ``--box``: rank 0 hosts the coordinator role,
every rank is a worker, and the broadcast / FedAvg legs run as fused NVLink kernels
(``colearn_federated_learning_b200.parallel``).
"""
import argparse
import logging
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from colearn_federated_learning_b200.control.arguments import Arguments  # noqa: E402
from colearn_federated_learning_b200.models import MODEL_REGISTRY  # noqa: E402


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Run Federated coordinator")
    # --- reference flags (fc.py:64-90) ---------------------------------------------------------
    parser.add_argument("--port", "-p", type=int, default=1883, help="port number of the where the broker is listining (default 1883)")
    parser.add_argument("--host", type=str, default="localhost", help="broker ip address (default localhost)")
    parser.add_argument("--topic", "-t", type=str, required=True, help="topic where the event must be published")
    parser.add_argument("--remote", "-r", action="store_true", help="Remote learning activation")
    parser.add_argument("--window", "-w", type=int, default=1, help="temporal window size (default 1)")
    parser.add_argument("--encryption", "-e", action="store_true", help="Simulates the encryption on two virtual workes")
    parser.add_argument("--federated_round", "-f", type=int, default=1, help="number of federated rounds (round > 1 trains 1000 batches per round)")
    parser.add_argument("--iot", "-i", action="store_true", help="enable iot validation (IP allow-list)")
    # --- everything the reference hard-codes in Arguments / Coordinator.__init__ ------------------
    d = Arguments()
    parser.add_argument("--model", choices=sorted(MODEL_REGISTRY), default=d.model)
    parser.add_argument("--loss", default=d.loss, choices=["auto", "bce", "sse", "xent", "mse"])
    parser.add_argument("--batch-size", type=int, default=d.batch_size)
    parser.add_argument("--local-epochs", type=int, default=d.epochs)
    parser.add_argument("--max-batches", type=int, default=d.federate_after_n_batches, help="local SGD steps per round (-1 = full epoch; rounds>1 default to 1000)")
    parser.add_argument("--lr", type=float, default=d.lr)
    parser.add_argument("--server-lr", type=float, default=d.server_lr)
    parser.add_argument("--seed", type=int, default=d.seed)
    parser.add_argument("--log-interval", type=int, default=d.log_interval)
    parser.add_argument("--test-path", type=str, default=d.test_path, help="CSV used by local/encrypted mode and evaluation")
    parser.add_argument("--synthetic", type=int, default=0, help="use N synthetic UNSW-shaped rows instead of --test-path")
    parser.add_argument("--dataset", default="unsw", help="dataset adapter that reads --test-path (data.register_dataset; default: the Bot-IoT CSV adapter)")
    parser.add_argument("--plugin", action="append", default=[], metavar="MODULE_OR_FILE",
                        help="import this module / .py file first: it may call models.register_model / data.register_dataset "
                             "(the reference's 'use your own model/dataset' without editing sources; repeatable)")
    parser.add_argument("--weighted", action="store_true", help="sample-count weighted FedAvg (default: uniform, like the reference)")
    parser.add_argument("--select", type=int, default=None, help="train at most k of the collected workers")
    parser.add_argument("--selection", choices=["all", "first", "random"], default="all")
    parser.add_argument("--checkpoint", type=str, default="./test.pth")
    parser.add_argument("--save-every", type=int, default=0, help="also write the checkpoint every N rounds (0 = only after the last round, like the reference)")
    parser.add_argument("--dtype", choices=["fp32", "bf16"], default=d.dtype,
                        help="--box: bf16 makes the FedAvg broadcast carry a bf16 copy of the model next to the fp32 master "
                             "(consumed directly by the tcgen05 GEMMs of the wide / conv models)")
    parser.add_argument("--no-cuda", action="store_true")
    parser.add_argument("--strict-events", action="store_true", help="validate IPv4 octets / full match (reference is lax)")
    parser.add_argument("--fit-timeout", type=float, default=None, help="seconds before a silent remote worker is dropped from a round")
    parser.add_argument("--filter-file", type=str, default=None)
    parser.add_argument("--metrics", type=str, default=None, help="write per-round JSONL here")
    parser.add_argument("--round-log", type=str, default=None, help="write the reference-format round log here")
    parser.add_argument("--metrics-port", type=int, default=0, help="serve Prometheus metrics (rounds, round time, losses, traffic) on this port")
    parser.add_argument("--evaluate", action="store_true", help="evaluate the global model on --test-path after training")
    parser.add_argument("--embedded-broker", action="store_true", help="start the TCP bus broker inside this process")
    parser.add_argument("--tls-ca", default=None, help="CA bundle: verify the broker / the devices (and, with --tls-cert on the embedded broker, demand client certificates)")
    parser.add_argument("--tls-cert", default=None, help="certificate for the embedded broker / client certificate towards broker and devices")
    parser.add_argument("--tls-key", default=None, help="private key belonging to --tls-cert")
    parser.add_argument("--tls-no-verify-hostname", action="store_true",
                        help="verify the certificate chain only, not that the certificate names the broker / device address")
    parser.add_argument("--inject", action="append", default=[], metavar="SPEC",
                        help="fault injection on the embedded broker: drop:<regex> | dup:<regex> | delay:<seconds>:<regex> (repeatable)")
    parser.add_argument("--exit-after", type=int, default=0, help="exit after N completed trainings (0 = run forever)")
    parser.add_argument("--box", action="store_true", help="multi-GPU box mode under torchrun (rank 0 = coordinator)")
    parser.add_argument("--backend", choices=["auto", "fused", "nccl", "cpu"], default="auto")
    parser.add_argument("--clients-per-gpu", type=int, default=1, help="--box: virtual federated devices hosted by each GPU (one CTA each)")
    parser.add_argument("--enc-items", type=int, default=None, help="-e: number of samples that are secret-shared and trained on (reference: 1000)")
    parser.add_argument("--round-deadline-ms", type=float, default=0.0,
                        help="--box: a selected worker that has not delivered its model this many ms after the coordinator started its "
                             "reduce is dropped from that round (weights renormalised); 0 = wait for everyone")
    parser.add_argument("--box-event", default="TRAINING", help="--box: the state every rank announces at start (rw.py --event)")
    parser.add_argument("--box-script", default=None, metavar="RANK:STATE:AFTER[,...]",
                        help="--box: further events, e.g. '3:NOT_READY:0,2:INFERENCE:1' = rank 3 withdraws right after its announcement, "
                             "rank 2 asks for inference once it has seen one training complete")
    parser.add_argument("--box-no-reannounce", action="store_true",
                        help="--box: devices do not ask for training again after a training they took part in")
    parser.add_argument("--box-inference-rows", type=int, default=5, help="--box: rows of its shard every rank tags as inference data")
    return parser


def arguments_from_cli(ns: argparse.Namespace) -> Arguments:
    a = Arguments()
    a.model, a.loss, a.batch_size, a.epochs = ns.model, ns.loss, ns.batch_size, ns.local_epochs
    a.federate_after_n_batches, a.lr, a.server_lr, a.seed = ns.max_batches, ns.lr, ns.server_lr, ns.seed
    a.log_interval, a.test_path, a.synthetic, a.weighted = ns.log_interval, ns.test_path, ns.synthetic, ns.weighted
    a.no_cuda, a.backend = ns.no_cuda, ns.backend
    if getattr(ns, "enc_items", None):
        a.n_train_items_enc = int(ns.enc_items)
    a.dtype, a.save_every = ns.dtype, max(0, ns.save_every)
    a.dataset = ns.dataset
    return a


def main(args: argparse.Namespace) -> None:
    logging.basicConfig(format="%(asctime)s: %(message)s", level=logging.INFO, datefmt="%H:%M:%S")
    logging.info(os.getpid())
    if args.box:
        from colearn_federated_learning_b200.parallel.box import run_box_coordinator
        run_box_coordinator(args, arguments_from_cli(args))
        return

    from colearn_federated_learning_b200.control.bus import TcpBroker
    from colearn_federated_learning_b200.control.coordinator import Coordinator
    from colearn_federated_learning_b200.utils.metrics import RoundLogger
    import socket
    import time

    from colearn_federated_learning_b200.control.tls import contexts_from_cli

    broker = None
    client_tls = contexts_from_cli(args.tls_ca, args.tls_cert, args.tls_key, server=False,
                                   check_hostname=not args.tls_no_verify_hostname)
    if args.embedded_broker:
        server_tls = contexts_from_cli(args.tls_ca, args.tls_cert, args.tls_key, server=True,
                                       require_client_cert=bool(args.tls_ca)) if args.tls_cert else None
        broker = TcpBroker("127.0.0.1" if args.host == "localhost" else args.host, args.port, ssl_context=server_tls).start()
        logging.info("embedded bus broker listening on %s:%d", broker.host, broker.port)
        for spec in args.inject:
            broker.broker.inject_from_spec(spec)
            logging.info("fault injection active: %s", spec)
    exporter = None
    if args.metrics_port:
        from colearn_federated_learning_b200.utils.metrics import PrometheusExporter
        exporter = PrometheusExporter(args.metrics_port)
        logging.info("Prometheus metrics on http://127.0.0.1:%d/metrics", exporter.port)
    coordinator = Coordinator(args.window, args.remote, args.federated_round, args.encryption, args.iot,
                              args=arguments_from_cli(args), transport="tcp", path=args.checkpoint,
                              strict_events=args.strict_events, select_k=args.select, selection=args.selection,
                              fit_timeout=args.fit_timeout, filter_file=args.filter_file,
                              metrics=RoundLogger(args.metrics, args.round_log, exporter=exporter), evaluate_after=args.evaluate,
                              worker_ssl_context=client_tls)
    if client_tls is not None:
        coordinator.tls_set(context=client_tls)
    try:
        if args.exit_after > 0:
            coordinator.run(args.host, args.port, args.topic, forever=False)
            # a window whose training raised counts too (it is in the scheduler's history with an "error" entry):
            # otherwise one failing training would keep this loop spinning forever
            def finished() -> int:
                failed = sum(1 for rec in coordinator.windower.history if "error" in rec)
                return coordinator.trainings_done + failed
            while finished() < args.exit_after:
                time.sleep(0.05)
            if coordinator.windower.last_error is not None:
                logging.error("last training failed: %r", coordinator.windower.last_error)
            coordinator.shutdown()
        else:
            coordinator.run(args.host, args.port, args.topic)
    except (ConnectionRefusedError, socket.gaierror) as e:
        logging.error("cannot reach the bus broker at %s:%d (%r); start one with --embedded-broker", args.host, args.port, e)
        sys.exit(2)
    finally:
        if broker is not None:
            broker.stop()


def parse_cli(argv=None) -> argparse.Namespace:
    """Plugins are imported before the real parser is built, so that ``--model`` accepts what they register."""
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--plugin", action="append", default=[])
    known, _ = pre.parse_known_args(argv)
    if known.plugin:
        from colearn_federated_learning_b200.models import load_plugins
        load_plugins(known.plugin)
    return build_parser().parse_args(argv)


if __name__ == "__main__":
    main(parse_cli())
