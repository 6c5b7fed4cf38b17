mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/rD_pytest_gpu.log 2>&1; echo "pytest -x rc=$?"; tail -n 3 gpurun_out/rD_pytest_gpu.log | cut -c1-250
timeout 200 python scripts/round_overhead.py --model ffnn --out gpurun_out/rD_round_overhead_ffnn_n1.json > gpurun_out/rD_round_overhead_ffnn_n1.log 2>&1; echo "overhead rc=$?"; tail -n 5 gpurun_out/rD_round_overhead_ffnn_n1.log | cut -c1-400
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/rD_launches_cfg2_small.csv python bench.py --config cfg2 --samples 1024 --steps 6 --warmup 3 --no-e2e > gpurun_out/rD_launches_cfg2_small.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/rD_launches_cfg2_small.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")) / 1e3)
    except ValueError: pass
for k, v in agg.items():
    print("%4d x  median %8.2f us  min %8.2f  max %8.2f  %s" % (len(v), sorted(v)[len(v)//2], min(v), max(v), k))
PY
timeout 200 python scripts/bench_convnet.py --only native_eager,native_graph > gpurun_out/rD_convnet_step.json 2> gpurun_out/rD_convnet_step.log; echo "convnet rc=$?"; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/rD_convnet_step.json') if l.startswith('{')][-1]; print('convnet graph ms/step', d['native_graph']['ms_per_step'], 'eager', d['native_eager']['ms_per_step'], 'launches', d['native_eager'].get('launches_per_fit'))"
timeout 200 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/rD_bench_cfg4_n1.json 2> gpurun_out/rD_bench_cfg4_n1.err; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/rD_bench_cfg4_n1.json') if l.startswith('{')][-1]; print('cfg4 n1', d['value'], d['e2e']['value'], d['config'].get('train_path'))"
timeout 200 python bench.py --config cfg5 --steps 8 --warmup 3 > gpurun_out/rD_bench_cfg5_n1.json 2> gpurun_out/rD_bench_cfg5_n1.err; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/rD_bench_cfg5_n1.json') if l.startswith('{')][-1]; print('cfg5 n1', d['value'], d['e2e']['value'], d['config'].get('roofline'))"
