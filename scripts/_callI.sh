mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -f -o gpurun_out/rI_prof_ffnn_sse python scripts/prof_mlp_only.py 8192 ffnn > gpurun_out/rI_prof_ffnn_sse.log 2>&1; echo "ncu ffnn rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -f -o gpurun_out/rI_prof_mlp64 python scripts/prof_mlp_only.py 8192 mlp > gpurun_out/rI_prof_mlp64.log 2>&1; echo "ncu mlp64 rc=$?"
