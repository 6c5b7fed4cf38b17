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
    on_gpu = backend != "cpu" and device.type == "cuda"
    engine = FederatedEngine(args.model, backend=backend, device=device, batch_size=args.batch_size, lr=args.lr,
                             local_epochs=args.epochs, max_batches=max_batches, loss=args.loss, weighted=args.weighted,
                             server_lr=args.server_lr, seed=args.seed,
                             bf16_shadow=on_gpu and getattr(args, "dtype", "fp32") == "bf16",
                             clients_per_rank=getattr(cli, "clients_per_gpu", 1) if on_gpu else 1)
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
    save_every = max(0, getattr(args, "save_every", 0))
    metrics_path = getattr(cli, "metrics", None)
    if metrics_path or save_every:
        # observable mode: one engine call per round, so every round gets its own device time (max over ranks),
        # traffic / link-roofline figures and, if asked for, a checkpoint.  Without these flags all rounds are
        # enqueued in one call and the host never enters the loop.
        _run_rounds_observed(engine, plan["mask"], rounds, rank, world, metrics_path, save_every, cli.checkpoint)
    else:
        rep = engine.run_rounds(rounds, masks=plan["mask"])
        if rank == 0:
            for i in range(rounds):
                for k in range(world):
                    if (plan["mask"] >> k) & 1:
                        log.info("Loss for worker id: %s tensor(%.4f)", "%s:%d" % rank_identity(k), float(rep.losses[i, k, 0]))
            log.info("Total training time: %s (device %.3f ms for %d rounds)", time.time() - t0, rep.device_ms, rounds)
    if rank == 0:
        engine.save_checkpoint(cli.checkpoint)
    shutdown()


NVLINK_GBPS_PER_DIR = 900.0     # NVLink 5 per GPU and direction (the roofline the box-mode metrics refer to)


def round_record(round_idx: int, mask: int, world: int, counts: List[int], losses, device_ms: float, bytes_bcast: int,
                 bytes_reduce: int, launches: int, algo: str) -> Dict[str, Any]:
    """One JSONL record per round (SURVEY §5 metrics): who trained, on how many samples, the losses, the device time
    (already the max over ranks), the model bytes moved by the two collective legs and how far the round is from the
    time the slower leg would take at NVLink line rate."""
    sel = [k for k in range(world) if (mask >> k) & 1]
    t = max(device_ms, 1e-6) * 1e-3
    link_s = max(bytes_bcast, bytes_reduce) / (NVLINK_GBPS_PER_DIR * 1e9)
    return {"round": round_idx, "selected": ["%s:%d" % rank_identity(k) for k in sel], "n_k": [counts[k] for k in sel],
            "loss_k": [float(losses[k]) for k in sel], "t_round_max_over_ranks_ms": device_ms, "algo": algo,
            "bytes_bcast": bytes_bcast, "bytes_reduce": bytes_reduce, "launches": launches,
            "GBps": (bytes_bcast + bytes_reduce) / t / 1e9, "link_time_lower_bound_ms": link_s * 1e3,
            "roofline_frac": link_s / t}


def _run_rounds_observed(engine: FederatedEngine, mask: int, rounds: int, rank: int, world: int, metrics_path: Optional[str],
                         save_every: int, checkpoint: str) -> None:
    import json

    import torch

    t0 = time.time()
    total_ms = 0.0
    for i in range(rounds):
        rep = engine.run_rounds(1, masks=mask)
        ms = torch.tensor([rep.device_ms], dtype=torch.float64, device=engine.device if engine.backend != "cpu" else "cpu")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms += float(ms)
        if rank == 0:
            losses = rep.losses[0, :, 0].tolist() if rep.losses is not None else [0.0] * world
            rec = round_record(i, mask, world, engine.counts, losses, float(ms), rep.bytes_bcast, rep.bytes_reduce,
                               rep.launches, rep.algo)
            for k, l in zip([k for k in range(world) if (mask >> k) & 1], rec["loss_k"]):
                log.info("Loss for worker id: %s tensor(%.4f)", "%s:%d" % rank_identity(k), l)
            log.info("Time round %d : %s", i, rec["t_round_max_over_ranks_ms"] * 1e-3)
            if metrics_path:
                with open(metrics_path, "a") as f:
                    f.write(json.dumps(rec) + "\n")
            if save_every and (i + 1) % save_every == 0 and i + 1 < rounds:
                engine.save_checkpoint(checkpoint)
                log.info("checkpoint after round %d written to %s", i + 1, checkpoint)
    if rank == 0:
        log.info("Total training time: %s (device %.3f ms for %d rounds)", time.time() - t0, total_ms, rounds)
