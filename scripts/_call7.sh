mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu.py tests/test_gpu_schedules.py -m gpu -q --timeout 280 -p no:cacheprovider -k "multi_rank or overlapped or never_signals or server_learning_rate or resnet_round" > gpurun_out/r2c7_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -n 14 gpurun_out/r2c7_pytest_multi.log | cut -c1-250
timeout 200 $TR --nproc-per-node 2 --master-port 29801 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c7_bench_reference_n2.json 2> gpurun_out/r2c7_bench_reference_n2.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/r2c7_bench_reference_n2.json
timeout 200 $TR --nproc-per-node 2 --master-port 29802 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c7_bench_default_n2.json 2> gpurun_out/r2c7_bench_default_n2.err; echo "ours rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c7_bench_default_n2.json"))
    print("ref_local n2", d["value"], d["e2e"]["value"], d["e2e"].get("per_call_sync_value"), d["gpu_launches"], d["config"]["self_check"])
    print("cfg2", {k:(v if k!="e2e" else v["value"]) for k,v in d["config"]["also_measured"]["cfg2"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2c7_bench_default_n2.err").read()[-1500:])
PY
timeout 200 $TR --nproc-per-node 2 --master-port 29803 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg5 > gpurun_out/r2c7_bench_cfg5_n2.json 2> gpurun_out/r2c7_bench_cfg5_n2.err; cut -c1-200 gpurun_out/r2c7_bench_cfg5_n2.json; python -c "import json;d=json.load(open('gpurun_out/r2c7_bench_cfg5_n2.json'));print(d['config'].get('roofline'), d['config'].get('nvls'), d['config'].get('train_path'))"
timeout 200 $TR --nproc-per-node 2 --master-port 29804 federated_coordinator.py -t topic/state --box --model ffnn --synthetic 4096 -w 1 --checkpoint gpurun_out/r2c7_box.pth --exit-after 3 --evaluate --box-script "1:NOT_READY:1,1:INFERENCE:1,1:TRAINING:2" -f 3 --metrics gpurun_out/r2c7_box_rounds.jsonl > gpurun_out/r2c7_box.log 2>&1; echo "box rc=$?"; grep -E "window closed|inference on|Loss evaluation|Total training" gpurun_out/r2c7_box.log | cut -c1-200
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/r2c7_microbench_mlp.json > gpurun_out/r2c7_microbench_mlp.log 2>&1; grep "'batch': 1, 'samples': 8192" gpurun_out/r2c7_microbench_mlp.log | cut -c40-175
