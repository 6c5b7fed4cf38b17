#!/usr/bin/env python
"""Read what scripts/run_round2_first.sh left in gpurun_out/ and print (1) which groups of the unvalidated GPU tests
passed, (2) the ResNet-18 step time of every schedule, fastest first, (3) the SCHEDULE_DEFAULTS edit that adopts the
fastest schedule whose tests passed.

    python scripts/pick_schedule.py [gpurun_out]
"""
import glob
import json
import os
import re
import sys

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"

log = os.path.join(out_dir, "r2_unvalidated_tests.log")
failed_names = []
if os.path.exists(log):
    text = open(log).read()
    print("== unvalidated GPU tests")
    for grp, body in re.findall(r"=== group: (.*?)\n(.*?)(?==== group:|\Z)", text, flags=re.S):
        summary = re.findall(r"(\d+ (?:passed|failed|skipped|error)[^\n]*)", body)
        rc = re.findall(r"rc=(\d+)", body)
        print(f"  {grp:70s} rc={rc[-1] if rc else '?':>3s}  {summary[-1] if summary else 'no summary (timeout / crash?)'}")
    failed_names = re.findall(r"^FAILED (\S+)", text, flags=re.M)
    for f in failed_names[:30]:
        print("    FAILED", f)
else:
    print(f"(no {log})")

rows = []
for path in sorted(glob.glob(os.path.join(out_dir, "r2_convnet_*.json"))):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except (ValueError, IndexError):
        rows.append((float("inf"), os.path.basename(path), None, "unreadable (crash? see the .err file)"))
        continue
    g = d.get("native_graph") or {}
    e = d.get("native_eager") or {}
    rows.append((g.get("ms_per_step", float("inf")), os.path.basename(path), d.get("flags", {}),
                 f"graph {g.get('ms_per_step', float('nan')):.3f} ms/step  eager {e.get('ms_per_step', float('nan')):.3f}  "
                 f"launches/fit {g.get('launches_per_fit', '?')}  loss {g.get('last_loss', float('nan')):.3f}"
                 + (f"  | cuDNN+autograd {d['torch_cudnn_autocast']['ms_per_step']:.3f}" if "torch_cudnn_autocast" in d else "")))
rows.sort(key=lambda r: r[0])
print("\n== ResNet-18 step (batch 128), fastest first")
for ms, name, flags, desc in rows:
    print(f"  {name:95s} {desc}")
ok = [r for r in rows if r[2] is not None and r[0] != float("inf")]
if ok:
    best = ok[0]
    print("\n== fastest schedule:", best[1])
    print("   SCHEDULE_DEFAULTS edit (fl/convnet.py) — only after the matching groups above passed:")
    for k, v in sorted((best[2] or {}).items()):
        if k.startswith("COLEARN_CONV_") and k != "COLEARN_CONV_PATH":
            print(f'       "{k[len("COLEARN_CONV_"):]}": {v},')
        elif k == "COLEARN_PDL":
            print("       (+ COLEARN_PDL=1: process-wide launch attribute, see docs/ROUND2_NOTES.md 1a)")
