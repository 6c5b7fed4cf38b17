#!/bin/sh
# Capture the evidence the judge asks for (run under gpurun, 1 GPU):
#   * SASS listing of every kernel (tcgen05 -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM, multimem.*)
#   * per-launch device times of one bench run
#   * ncu --set full captures of the top kernels (persistent MLP, tcgen05 GEMM)
# Outputs land in gpurun_out/; copy the summaries you want kept into profiles/.
set -u
mkdir -p gpurun_out
SO=$(ls colearn_federated_learning_b200/ops/_colearn_C*.so | head -n 1)
cuobjdump -sass "$SO" | grep -E "Function|UTCHMMA|UTMALDG|LDTM|UTCBAR|MULTIMEM|SYNCS" > gpurun_out/sass_mnemonics.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd -s 2 -c 1 -o gpurun_out/prof_mlp \
    python scripts/microbench.py --only mlp --out gpurun_out/mb_prof.json > gpurun_out/prof_mlp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 1 -o gpurun_out/prof_gemm \
    python scripts/microbench.py --only gemm --quick --out gpurun_out/mb_prof_gemm.json > gpurun_out/prof_gemm.log 2>&1
