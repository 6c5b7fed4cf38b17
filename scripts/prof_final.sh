#!/bin/sh
# ncu --set full captures of the hot kernels (1 GPU): GEMM (2SM / 1-CTA / wgrad epilogue), persistent MLP, FedAvg, SGD
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -s 6 -c 3 -o gpurun_out/prof_gemm python scripts/prof_gemm_only.py 4096 > gpurun_out/prof_gemm.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:"fedavg_kernel|sgd_step_kernel" -s 4 -c 2 -o gpurun_out/prof_elem python scripts/prof_gemm_only.py 1024 > gpurun_out/prof_elem.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -o gpurun_out/prof_mlp_final python scripts/prof_mlp_only.py 8192 > gpurun_out/prof_mlp_final.log 2>&1
tail -n 2 gpurun_out/prof_gemm.log gpurun_out/prof_elem.log gpurun_out/prof_mlp_final.log | cut -c1-200
