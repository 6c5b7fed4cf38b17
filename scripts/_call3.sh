mkdir -p gpurun_out
# 1. headline kernel: variants 5 (default) / 6 (packed FFMA2) / 3 / 1
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/r2c3_microbench_mlp.json > gpurun_out/r2c3_microbench_mlp.log 2>&1
grep variant gpurun_out/r2c3_microbench_mlp.log | cut -c1-200
timeout 100 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "persistent_mlp" -p no:cacheprovider 2>&1 | tail -n 3
# 2. fused wgrad -> reduce: which chunks are never published?
COLEARN_OVERLAP_TIMEOUT_S=4 timeout 100 python scripts/debug_overlap.py 512 3 128 > gpurun_out/r2c3_debug_overlap_small.json 2> gpurun_out/r2c3_debug_overlap_small.err; echo "overlap small rc=$?"; cut -c1-1500 gpurun_out/r2c3_debug_overlap_small.json
COLEARN_OVERLAP_TIMEOUT_S=4 timeout 100 python scripts/debug_overlap.py 4096 4 1024 > gpurun_out/r2c3_debug_overlap_wide.json 2> gpurun_out/r2c3_debug_overlap_wide.err; echo "overlap wide rc=$?"; cut -c1-1500 gpurun_out/r2c3_debug_overlap_wide.json
# 3. split-K: rounding-sized perturbation control
timeout 200 python scripts/debug_splitk_step.py COLEARN_CONV_SPLITK=1 > gpurun_out/r2c3_debug_splitk.json 2> gpurun_out/r2c3_debug_splitk.err; echo "splitk rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c3_debug_splitk.json"))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!="worst"}) for k,v in d.items()})
PY
# 4. cfg5 at N=1: operands in place
for flags in "" "COLEARN_MLP_DGRAD_KN=1" "COLEARN_MLP_DGRAD_KN=1 COLEARN_MLP_WGRAD_MN=1"; do
  tag=$(echo "x$flags" | tr ' =' '__')
  env $flags timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > "gpurun_out/r2c3_bench_cfg5_${tag}.json" 2> "gpurun_out/r2c3_bench_cfg5_${tag}.err"
  echo "cfg5 $flags: $(cut -c1-200 gpurun_out/r2c3_bench_cfg5_${tag}.json)"
done
# 5. ResNet step: new defaults, + PDL, SPLITK=2, legacy
for flags in "" "COLEARN_PDL=1" "COLEARN_CONV_SPLITK=2" "COLEARN_CONV_SPLITK=2 COLEARN_PDL=1" "COLEARN_CONV_STREAMS=0" ; do
  tag=$(echo "x$flags" | tr ' =' '__')
  env $flags timeout 60 python scripts/bench_convnet.py --reps 3 --only native_eager,native_graph > "gpurun_out/r2c3_convnet_${tag}.json" 2> "gpurun_out/r2c3_convnet_${tag}.err"
  echo "conv $flags: $(cut -c1-420 gpurun_out/r2c3_convnet_${tag}.json)"
done
timeout 100 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/r2c3_bench_cfg4_n1.json 2> gpurun_out/r2c3_bench_cfg4_n1.err; cut -c1-300 gpurun_out/r2c3_bench_cfg4_n1.json
# 6. ncu of the headline kernel (variant 5 = default, then 6)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -o gpurun_out/r2c3_prof_mlp_v5 python scripts/prof_mlp_only.py 8192 > gpurun_out/r2c3_prof_mlp_v5.log 2>&1
COLEARN_MLP_VARIANT=6 timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_local_sgd_kernel_v2 -s 2 -c 1 -o gpurun_out/r2c3_prof_mlp_v6 python scripts/prof_mlp_only.py 8192 > gpurun_out/r2c3_prof_mlp_v6.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -n 3
