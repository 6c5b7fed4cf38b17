"""Where does a round's time go besides the SGD steps?  Fixed cost per round = intercept of (round time) over (steps per round).

    python scripts/round_overhead.py [--model ffnn] [--out gpurun_out/round_overhead.json]          (1 GPU, or under torchrun)

For n in a few shard sizes: K rounds in ONE run_rounds call (what a training does) and K single-round calls (what bench.py's
device-timed value sums), CUDA events, max over ranks; then the per-phase event timers of the engine (bcast / local_fit /
reduce_apply) for single-round calls.  The least-squares line through (n, ms per round) gives us per step and the fixed cost.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from colearn_federated_learning_b200.data import synthetic_unsw
from colearn_federated_learning_b200.parallel import FederatedEngine, init_distributed, shutdown


def fit_line(pts):
    n = len(pts)
    sx, sy = sum(p[0] for p in pts), sum(p[1] for p in pts)
    sxx, sxy = sum(p[0] * p[0] for p in pts), sum(p[0] * p[1] for p in pts)
    slope = (n * sxy - sx * sy) / (n * sxx - sx * sx)
    return slope, (sy - slope * sx) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ffnn")
    ap.add_argument("--loss", default="sse")
    ap.add_argument("--sizes", default="256,1024,4096,8192")
    ap.add_argument("--rounds", type=int, default=24)
    ap.add_argument("--out", default="gpurun_out/round_overhead.json")
    args = ap.parse_args()
    rank, world, device = init_distributed()

    def mx(v):
        t = torch.tensor([v], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows = []
    K = args.rounds
    for n in [int(s) for s in args.sizes.split(",")]:
        x, y = synthetic_unsw(n, seed=3 + rank)
        eng = FederatedEngine(args.model, backend="fused", device=device, batch_size=1, lr=0.01, local_epochs=1, weighted=False,
                              loss=args.loss if args.model == "ffnn" else "auto", seed=1)
        eng.set_local_data(x, y)
        for _ in range(3):
            eng.run_rounds(1)
        eng.run_rounds(4)
        multi = min(mx(eng.run_rounds(K).device_ms) for _ in range(3)) / K
        single = []
        for _ in range(3):
            t = 0.0
            for _ in range(K):
                t += eng.run_rounds(1).device_ms
            single.append(mx(t) / K)
        eng.phase_timing = True
        ph = None
        for _ in range(4):
            ph = eng.run_rounds(1).extra.get("phases_ms")
        eng.phase_timing = False
        rows.append({"n": n, "ms_per_round_one_call": multi, "ms_per_round_single_calls": min(single), "phases_ms_single_call_rank0": ph})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
        eng.close() if hasattr(eng, "close") else None
    out = {"model": args.model, "world": world, "rounds": K, "rows": rows}
    for key in ("ms_per_round_one_call", "ms_per_round_single_calls"):
        s, o = fit_line([(r["n"], r[key]) for r in rows])
        out[key + "_fit"] = {"us_per_step": s * 1e3, "fixed_us_per_round": o * 1e3}
    if rank == 0:
        print(json.dumps({k: v for k, v in out.items() if k.endswith("_fit")}), flush=True)
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    shutdown()


if __name__ == "__main__":
    main()
