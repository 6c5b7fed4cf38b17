mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29711 scripts/comm_sweep.py --sizes 11181644,50397188 --nvlink-counters --out gpurun_out/rC_nvlink_twoshot_nvls.json > gpurun_out/rC_nvlink_twoshot_nvls.log 2>&1; echo "sweep nvls rc=$?"
timeout 200 $TR --master-port 29712 scripts/comm_sweep.py --sizes 11181644,50397188 --nvls 0 --nvlink-counters --out gpurun_out/rC_nvlink_twoshot_p2p.json > gpurun_out/rC_nvlink_twoshot_p2p.log 2>&1; echo "sweep p2p rc=$?"
timeout 200 $TR --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 3 --config cfg2 --nvlink-counters > gpurun_out/rC_bench_cfg2_n2_nvlink.json 2> gpurun_out/rC_bench_cfg2_n2_nvlink.err; echo "cfg2 rc=$?"
timeout 200 $TR --master-port 29714 bench.py --gpus 2 --steps 8 --warmup 3 --config cfg5 --nvlink-counters > gpurun_out/rC_bench_cfg5_n2_nvlink.json 2> gpurun_out/rC_bench_cfg5_n2_nvlink.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for f in ("rC_nvlink_twoshot_nvls", "rC_nvlink_twoshot_p2p"):
    try:
        for r in json.load(open(f"gpurun_out/{f}.json")):
            print(f, r["P"], "ms", round(r["twoshot_ms"], 4), "busbw", round(r["twoshot_busbw_GBps"], 1), "nvlink", json.dumps(r.get("nvlink"))[:600])
    except Exception as e:
        print(f, "failed", e)
for f in ("rC_bench_cfg2_n2_nvlink", "rC_bench_cfg5_n2_nvlink"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        print(f, d["value"], json.dumps(d["config"].get("nvlink"))[:600])
    except Exception as e:
        print(f, "failed", e)
PY
