#!/bin/sh
# What the driver runs at round end, on one GPU: pytest -m gpu -x, smoke(), both bench arms at N = 1; then the microbench and the
# ncu captures of the conv step and of the MLP kernels.  -> gpurun_out/rB_*
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/rB_pytest_gpu.log 2>&1; echo "pytest -x rc=$?"; tail -n 6 gpurun_out/rB_pytest_gpu.log | cut -c1-250
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rB_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/rB_smoke.log | cut -c1-200
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/rB_bench_reference_n1.json 2> gpurun_out/rB_bench_reference_n1.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/rB_bench_reference_n1.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/rB_bench_default_n1.json 2> gpurun_out/rB_bench_default_n1.err; echo "ours rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/rB_bench_default_n1.json') if l.startswith('{')][-1]
print('ref_local n1', d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], '| cfg2', d['config']['also_measured']['cfg2']['value'], d['config']['also_measured']['cfg2']['e2e']['value'])"
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/rB_microbench_mlp.json > gpurun_out/rB_microbench_mlp.log 2>&1; grep "'batch': 1, 'samples': 8192" gpurun_out/rB_microbench_mlp.log | cut -c40-175
sh scripts/prof_conv.sh 2>&1 | tail -n 20 | cut -c1-200
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -f -o gpurun_out/rB_prof_mlp64_default python scripts/prof_mlp_only.py 8192 mlp > gpurun_out/rB_prof_mlp64_default.log 2>&1; echo "ncu mlp64 rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -f -o gpurun_out/rB_prof_ffnn_default python scripts/prof_mlp_only.py 8192 ffnn > gpurun_out/rB_prof_ffnn_default.log 2>&1; echo "ncu ffnn rc=$?"
