"""Per-instruction view of a kernel's hot loop from an ncu report captured with --import-source on:

    python scripts/ncu_hot_loop.py gpurun_out/x.ncu-rep [min_executions] [--all]

Prints the SASS instructions executed at least `min_executions` times (default: 90 % of the maximum) with their stall samples,
share of all samples and dominant stall reason, plus a one-line summary (cycles, instructions per step if steps known)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            data.append((int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), r[ix["Source"]],
                         {k: int(r[ix[k]] or 0) for k in keys}))
        except ValueError:
            pass
    tot = sum(d[0] for d in data)
    mx = max(d[1] for d in data)
    thr = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else int(0.2 * mx)
    hot = [d for d in data if d[1] >= thr]
    print(f"{rep}: {len(data)} instructions, {tot} samples; {len(hot)} instructions executed >= {thr} times hold {sum(d[0] for d in hot)} samples")
    agg = {}
    for d in hot:
        for k, v in d[3].items():
            agg[k] = agg.get(k, 0) + v
    print("stall mix of the hot instructions:", ", ".join(f"{k[6:]} {100 * v / max(1, sum(agg.values())):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for i, d in enumerate(data):
        if d[1] >= thr:
            top = max(d[3], key=d[3].get)
            print(f"{i:5d} {d[1]:7d} {d[0]:5d} {100 * d[0] / tot:5.2f}%  {top[6:]:16s} {d[2][:100]}")


if __name__ == "__main__":
    main()
