"""``federated_coordinator.py --box``: the reference's roles mapped onto the GPUs of one box.

Rank 0 runs the real :class:`~..control.coordinator` control plane pieces — in-process pub/sub
broker, event parser (+ allow-list), temporal window, selection — and every rank plays a device:
it announces ``(10.0.0.<rank+1>, <8777+rank>, TRAINING)`` through the process-group store, a relay
thread on rank 0 republishes those payloads on the in-process bus (this replaces MQTT),
the window collects them, and when it closes rank 0 turns the snapshot into a **round plan**
(selection bitmask, rounds, hyper-parameters) that is handed to all ranks.  Every rank then
enters ``FederatedEngine.run_rounds`` where broadcast / local SGD / FedAvg are fused NVLink
kernels; finally rank 0 writes ``test.pth``.
"""
from __future__ import annotations

import logging
import threading
import time
from typing import Any, Dict, List, Optional

import torch.distributed as dist

from ..control.arguments import Arguments, ROUND_MODE_BATCHES
from ..control.bus import BusClient, InProcessBroker
from ..control.event_parser import EventParser, format_event
from ..control.selection import SelectionPolicy, LOWER_BOUND, UPPER_BOUND
from ..control.window import TemporalWindow
from ..data import synthetic_for_model, NetworkTrafficDataset
from ..settings import DeviceRegistry
from ..utils.checkpoint import load_or_init, checkpoint_compatible
from ..models import flatten_params
from .engine import FederatedEngine
from .launcher import init_distributed, shutdown

log = logging.getLogger(__name__)

BASE_PORT = 8777


def rank_identity(rank: int):
    return f"10.0.0.{rank + 1}", BASE_PORT + rank


def worker_id_to_rank(worker_id: str) -> int:
    return int(worker_id.rsplit(":", 1)[1]) - BASE_PORT


def collect_plan(world: int, window_s: float, topic: str, payloads: List[str], iot: bool, select_k: Optional[int],
                 selection: str, seed: int, filter_file: Optional[str] = None, strict: bool = False) -> Dict[str, Any]:
    """Rank-0 control plane: publish the ranks' events on the in-process bus, run the temporal
    window, return the selection mask of the first window that fires."""
    broker = InProcessBroker()
    registry = DeviceRegistry()
    kwargs = {"filter_file": filter_file} if filter_file else {}
    parser = EventParser(iot, strict=strict, **kwargs)
    fired = threading.Event()
    plan: Dict[str, Any] = {}

    def train_fn(snapshot):
        sel = SelectionPolicy(LOWER_BOUND, UPPER_BOUND, selection, select_k, seed).select(snapshot)
        mask = 0
        for wid in sel:
            mask |= 1 << worker_id_to_rank(wid)
        plan.update(mask=mask, members=list(snapshot.keys()), selected=list(sel.keys()))
        for wid in sel:
            registry.remove(wid)
        fired.set()

    windower = TemporalWindow(registry, window_s, train_fn, lower_bound=0)

    class Sub(BusClient):
        def on_message(self, client, userdata, msg):
            ev = parser.parse(msg.payload, remote=True)
            if ev is None:
                log.info("Some problems occurred")
                return
            if ev.state == "TRAINING":
                windower.on_training(ev.worker_id, ev.worker_id)
            elif ev.state == "NOT_READY":
                windower.on_not_ready(ev.worker_id)

    sub = Sub("coordinator", broker=broker)
    sub.connect()
    sub.subscribe(topic)
    sub.loop_start()
    pub = BusClient("relay", broker=broker)
    pub.connect()
    for p in payloads:
        pub.publish(topic, p)
    if not fired.wait(timeout=window_s + 30):
        plan.update(mask=0, members=[], selected=[])
    sub.loop_stop()
    return plan


def run_box_coordinator(cli, args: Arguments) -> None:
    rank, world, device = init_distributed()
    # every rank announces itself (state from --event would be TRAINING); rank 0 gathers the payloads
    ip, port = rank_identity(rank)
    payload = format_event(ip, "TRAINING", port)
    payloads: List[Optional[str]] = [None] * world
    if world > 1:
        dist.all_gather_object(payloads, payload)
    else:
        payloads = [payload]
    plan_box: List[Any] = [None]
    if rank == 0:
        plan_box[0] = collect_plan(world, float(cli.window), cli.topic, [p for p in payloads if p], cli.iot, cli.select,
                                   cli.selection, args.seed, cli.filter_file, cli.strict_events)
        log.info("window closed: members=%s selected=%s", plan_box[0].get("members"), plan_box[0].get("selected"))
    if world > 1:
        dist.broadcast_object_list(plan_box, src=0)
    plan = plan_box[0]
    if not plan or plan["mask"] == 0:
        log.info("No behaviour defined for the number of devices achieved")
        shutdown()
        return
    rounds = max(1, cli.federated_round)
    max_batches = args.federate_after_n_batches
    if rounds > 1 and max_batches < 0:
        max_batches = ROUND_MODE_BATCHES
    backend = "auto" if args.backend in ("auto", "nccl") else args.backend
    engine = FederatedEngine(args.model, backend=backend, device=device, batch_size=args.batch_size, lr=args.lr,
                             local_epochs=args.epochs, max_batches=max_batches, loss=args.loss, weighted=args.weighted,
                             server_lr=args.server_lr, seed=args.seed,
                             clients_per_rank=getattr(cli, "clients_per_gpu", 1) if backend != "cpu" and device.type == "cuda" else 1)
    if rank == 0:
        import os
        if os.path.exists(cli.checkpoint) and checkpoint_compatible(engine.model, cli.checkpoint):
            load_or_init(engine.model, cli.checkpoint)
            engine.load_global(flatten_params(engine.model))
    # private shards: contiguous ceil(N/world) split of the dataset, like dataset.federate(workers)
    if args.synthetic and args.synthetic > 0:
        x, y = synthetic_for_model(args.model, args.synthetic, seed=args.seed)
    else:
        x, y = NetworkTrafficDataset(args.test_path).tensors()
    from ..data import shard_bounds
    lo, hi = shard_bounds(len(x), world)[rank]
    engine.set_local_data(x[lo:hi], y[lo:hi])
    t0 = time.time()
    rep = engine.run_rounds(rounds, masks=plan["mask"])
    if rank == 0:
        for i in range(rounds):
            for k in range(world):
                if (plan["mask"] >> k) & 1:
                    log.info("Loss for worker id: %s tensor(%.4f)", "%s:%d" % rank_identity(k), float(rep.losses[i, k, 0]))
        log.info("Total training time: %s (device %.3f ms for %d rounds)", time.time() - t0, rep.device_ms, rounds)
        engine.save_checkpoint(cli.checkpoint)
    shutdown()
