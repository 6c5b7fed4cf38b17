#!/bin/sh
# A/B of a change to the persistent MLP kernel on one GPU (~30 s): every variant's us/step, the kernel tests, the default bench line.
#   -> gpurun_out/rL_*
mkdir -p gpurun_out
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/rL_microbench_mlp.json > gpurun_out/rL_microbench_mlp.log 2>&1; grep "'batch': 1, 'samples': 8192" gpurun_out/rL_microbench_mlp.log | cut -c40-175
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider > gpurun_out/rL_pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 gpurun_out/rL_pytest_kernels.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/rL_bench_default_n1.json 2> gpurun_out/rL_bench_default_n1.err; echo "ours rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/rL_bench_default_n1.json') if l.startswith('{')][-1]
print('ref_local n1', d['value'], d['e2e']['value'], '| cfg2', d['config']['also_measured']['cfg2']['value'], d['config']['also_measured']['cfg2']['e2e']['value'])"
