mkdir -p gpurun_out
COLEARN_OVERLAP_TIMEOUT_S=5 timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2c5_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2c5_pytest_gpu.log | cut -c1-220
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/r2c5_microbench_mlp.json > gpurun_out/r2c5_microbench_mlp.log 2>&1; grep variant gpurun_out/r2c5_microbench_mlp.log | cut -c1-190
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c5_bench_default_n1.json 2> gpurun_out/r2c5_bench_default_n1.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c5_bench_default_n1.json"))
print("ref_local", d["value"], d["e2e"], d["gpu_launches"], d["config"]["pipelined_rounds_per_s"], d["config"]["self_check"])
print("cfg2", d["config"]["also_measured"])
PY
sh scripts/sanitize.sh memcheck
sh scripts/sanitize.sh racecheck
