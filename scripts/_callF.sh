mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 3 --nvlink-counters > gpurun_out/rF_bench_default_n2.json 2> gpurun_out/rF_bench_default_n2.err; echo "default rc=$?"
timeout 200 $TR --master-port 29732 bench.py --gpus 2 --steps 20 --warmup 3 --config cfg2 --nvlink-counters > gpurun_out/rF_bench_cfg2_n2_nvlink.json 2> gpurun_out/rF_bench_cfg2_n2_nvlink.err; echo "cfg2 rc=$?"
timeout 200 $TR --master-port 29734 scripts/round_overhead.py --model ffnn --sizes 128,512,2048,4096 --out gpurun_out/rF_round_overhead_ffnn_n2.json > gpurun_out/rF_round_overhead_ffnn_n2.log 2>&1; echo "overhead rc=$?"; tail -n 1 gpurun_out/rF_round_overhead_ffnn_n2.log | cut -c1-400
python - <<'PY'
import json
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        c = d["config"]
        print(f, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "pipelined", c.get("pipelined_rounds_per_s"), "check", (c.get("self_check") or {}).get("ok"),
              "nvls", c.get("nvls"), "nvlink", json.dumps(c.get("nvlink"))[:420])
    except Exception as e:
        print(f, "failed", e)
PY
