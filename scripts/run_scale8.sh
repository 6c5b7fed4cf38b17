#!/bin/sh
# One 8-GPU box (gpurun --gpus 8 -- 'sh scripts/run_scale8.sh'): the default bench line at N = 8 and N = 4 (with the NVLink counters),
# the fixed cost of a round at N = 8 (scripts/round_overhead.py) and BASELINE config 5 at N = 8 / 4.  -> gpurun_out/rM_*
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out/rM
timeout 150 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 3 --nvlink-counters > ${O}_default_n8.json 2> ${O}_default_n8.err; echo "default n8 rc=$?"
timeout 120 $TR --nproc-per-node 8 --master-port 29802 scripts/round_overhead.py --model ffnn --sizes 128,512,1024 --rounds 16 --out ${O}_round_overhead_ffnn_n8.json > ${O}_round_overhead_ffnn_n8.log 2>&1; echo "overhead rc=$?"; tail -n 1 ${O}_round_overhead_ffnn_n8.log | cut -c1-400
timeout 150 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 --steps 8 --warmup 3 --config cfg5 --nvlink-counters > ${O}_cfg5_n8.json 2> ${O}_cfg5_n8.err; echo "cfg5 n8 rc=$?"
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 150 $TR --nproc-per-node 4 --master-port 29804 bench.py --gpus 4 --steps 20 --warmup 3 > ${O}_default_n4.json 2> ${O}_default_n4.err ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 150 $TR --nproc-per-node 4 --master-port 29805 bench.py --gpus 4 --steps 8 --warmup 3 --config cfg5 > ${O}_cfg5_n4.json 2> ${O}_cfg5_n4.err ) &
wait
python - <<'PY'
import json
for f in ("rM_default_n8", "rM_default_n4", "rM_cfg5_n8", "rM_cfg5_n4"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        c = d["config"]
        a = c.get("also_measured", {}).get("cfg2")
        print(f, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "pipelined", round(c.get("pipelined_rounds_per_s") or 0, 2), "check", (c.get("self_check") or {}).get("ok"),
              "nvls", c.get("nvls"), "| cfg2", a and (round(a["value"], 2), round(a["e2e"]["value"], 2)), "roofline", (c.get("roofline") or {}).get("roofline_frac"))
        if c.get("nvlink"):
            print("   nvlink tx/round by rank", [int(v) for v in c["nvlink"]["tx_bytes_per_round_by_rank"]], "model", c["nvlink"]["model_bytes"])
    except Exception as e:
        print(f, "failed", repr(e))
PY
