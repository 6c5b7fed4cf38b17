#!/bin/sh
# First GPU calls of the next round, in stages that each fit one gpurun call (a call costs ~1.5 GPU-minutes before the command
# even starts; stage run times measured on the CPU box are not available, the estimates are from round 1's per-process costs):
#     sh scripts/run_round2_first.sh tests   (~10 min)  the opt-in GPU tests, one pytest process per group
#     sh scripts/run_round2_first.sh conv    (~9 min)   ResNet-18 step under every schedule
#     sh scripts/run_round2_first.sh prof    (~8 min)   per-kernel durations, ncu --set full of the conv kernels, memcheck, cfg4 / cfg1
#     sh scripts/run_round2_first.sh gemm    (~6 min)   cfg5 serial / overlapped / staged, GEMM epilogue bench, MLP kernel variants, cfg2
#     sh scripts/run_round2_first.sh         = all four
# What the stages do:
#   1. the GPU tests written after round 1's budget ran out (incl. the opt-in kernels / schedules); if the MN-major
#      GEMM tests fail, the descriptor sweep (scripts/debug_umma_mn.py) runs right away
#   2. the ResNet step with each opt-in re-scheduling, against the default and the cuDNN path
#   3. per-kernel durations of one eager ResNet step (where do the 1.77 ms go?) + ncu --set full of the conv kernels
#   4. compute-sanitizer memcheck over the conv ops
#   5. cfg4 at N=1 with the default path, cfg1 on a GPU
# Everything lands in gpurun_out/r2_*; copy the summaries into profiles/.
mkdir -p gpurun_out
STAGE=${1:-all}
want() { [ "$STAGE" = all ] || [ "$STAGE" = "$1" ]; }
if want tests; then
  # one pytest process per group: a kernel that hangs only takes its own group down (timeout kills the process + context)
  : > gpurun_out/r2_unvalidated_tests.log
  for grp in "resnet_eval" "fused_batchnorm or optional_step" "splitk_reduce or gemm_split_k or split_k_step" \
             "mn_major_operands or mn_major_b_operand or mn_major_fused" "mn_major_wgrad_step" \
             "implicit_conv_forward" "implicit_conv_wgrad or implicit_conv_dgrad_packed" "implicit_step" \
             "programmatic_dependent_launch" "overlapped_reduce" "staged_epilogue" "weights_in_place" "new_variants" "pipelined_read_back"; do
    echo "=== group: $grp" >> gpurun_out/r2_unvalidated_tests.log
    COLEARN_RUN_UNVALIDATED=1 timeout 240 python -m pytest tests/test_zz_round2_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "$grp" \
        >> gpurun_out/r2_unvalidated_tests.log 2>&1
    echo "rc=$? ($grp)" | tee -a gpurun_out/r2_unvalidated_tests.log
  done
  grep -E "passed|failed|error" gpurun_out/r2_unvalidated_tests.log | tail -n 12
  grep -E "^FAILED|^ERROR" gpurun_out/r2_unvalidated_tests.log | cut -c1-160 | head -n 40
  if grep -E "^FAILED.*(mn_major|b_operand)" gpurun_out/r2_unvalidated_tests.log > /dev/null; then
    timeout 300 python scripts/debug_umma_mn.py > gpurun_out/r2_umma_mn_sweep.log 2>&1
    cat gpurun_out/r2_umma_mn_sweep.log
  fi
fi
if want conv; then
  timeout 90 python scripts/bench_convnet.py --reps 3 > gpurun_out/r2_convnet_default.json 2> gpurun_out/r2_convnet_default.err
  echo "== default"; cut -c1-600 gpurun_out/r2_convnet_default.json
  ALL="COLEARN_CONV_WGRAD_MN=1 COLEARN_CONV_DGRAD_KN=1 COLEARN_CONV_SPLITK=1 COLEARN_CONV_FUSED_BN=1"
  IMP="$ALL COLEARN_CONV_IMPLICIT=2"
  for flags in "COLEARN_CONV_FUSED_BN=1" "COLEARN_CONV_STREAMS=1" "COLEARN_CONV_SHADOW_T=1" "COLEARN_CONV_SPLITK=1" "COLEARN_CONV_SPLITK=2" \
               "COLEARN_CONV_WGRAD_MN=1" "COLEARN_CONV_DGRAD_KN=1" "COLEARN_CONV_WGRAD_MN=1 COLEARN_CONV_SPLITK=1" \
               "COLEARN_CONV_FUSED_BN=1 COLEARN_CONV_STREAMS=1 COLEARN_CONV_SHADOW_T=1" \
               "COLEARN_CONV_IMPLICIT=1" "COLEARN_CONV_IMPLICIT=2" "$IMP" "$IMP COLEARN_CONV_STREAMS=1" \
               "$ALL" "$ALL COLEARN_CONV_STREAMS=1" "COLEARN_PDL=1" "$IMP COLEARN_PDL=1" "COLEARN_CONV_WGRAD_MN=1 COLEARN_CONV_DGRAD_KN=1 COLEARN_CONV_SPLITK=2 COLEARN_CONV_FUSED_BN=1"; do
    tag=$(echo "$flags" | tr ' =' '__' | sed 's/COLEARN_CONV_//g')
    env $flags timeout 60 python scripts/bench_convnet.py --reps 3 --only native_eager,native_graph \
        > "gpurun_out/r2_convnet_${tag}.json" 2> "gpurun_out/r2_convnet_${tag}.err"
    echo "== $tag rc=$?"; cut -c1-500 "gpurun_out/r2_convnet_${tag}.json"
  done
fi
if want prof; then
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_convnet_launch_times.csv \
      python scripts/prof_convnet_only.py 2 > gpurun_out/r2_convnet_launch_times.log 2>&1
  env $IMP timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_convnet_launch_times_allflags.csv \
      python scripts/prof_convnet_only.py 2 > gpurun_out/r2_convnet_launch_times_allflags.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"im2col_kernel|col2im_kernel|bn_reduce_kernel|bn_apply_kernel|bn_bwd_kernel" \
      -s 10 -c 10 -o gpurun_out/r2_prof_conv python scripts/prof_convnet_only.py 2 > gpurun_out/r2_prof_conv.log 2>&1
  timeout 200 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck_conv.log \
      python -m pytest tests/test_conv_ops.py -q -m gpu -x > gpurun_out/r2_sanitizer_memcheck_conv.out 2>&1
  echo "memcheck rc=$?"; tail -n 3 gpurun_out/r2_sanitizer_memcheck_conv.log
  timeout 80 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4_n1.json 2> gpurun_out/r2_bench_cfg4_n1.err
  cut -c1-300 gpurun_out/r2_bench_cfg4_n1.json
  timeout 60 python bench.py --config cfg1 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg1.json 2> gpurun_out/r2_bench_cfg1.err
  cut -c1-200 gpurun_out/r2_bench_cfg1.json
fi
if want gemm; then
  # cfg5 (wide MLP on the tcgen05 layer-wise trainer) at N=1: serial round vs fused wgrad -> FedAvg reduce (world 1 only shows the cost
  # of the reports + the capped grids; the gain needs N > 1: scripts/run_round2_8gpu.sh)
  timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1.json 2> gpurun_out/r2_bench_cfg5_n1.err
  COLEARN_OVERLAP_REDUCE=1 timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1_overlap.json 2> gpurun_out/r2_bench_cfg5_n1_overlap.err
  cut -c1-260 gpurun_out/r2_bench_cfg5_n1.json gpurun_out/r2_bench_cfg5_n1_overlap.json
  # row-per-thread vs line-coalesced GEMM epilogue per mode (wide-MLP layer shapes), then cfg5 with it
  timeout 120 python scripts/bench_gemm_epilogue.py > gpurun_out/r2_gemm_epilogue.json 2> gpurun_out/r2_gemm_epilogue.err
  cut -c1-1500 gpurun_out/r2_gemm_epilogue.json
  COLEARN_GEMM_STAGED=1 timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1_staged.json 2> gpurun_out/r2_bench_cfg5_n1_staged.err
  cut -c1-260 gpurun_out/r2_bench_cfg5_n1_staged.json
  COLEARN_MLP_DGRAD_KN=1 timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1_dgradkn.json 2> gpurun_out/r2_bench_cfg5_n1_dgradkn.err
  COLEARN_MLP_DGRAD_KN=1 COLEARN_GEMM_STAGED=1 timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1_dgradkn_staged.json 2> gpurun_out/r2_bench_cfg5_n1_dgradkn_staged.err
  COLEARN_MLP_DGRAD_KN=1 COLEARN_MLP_WGRAD_MN=1 COLEARN_GEMM_STAGED=1 timeout 80 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_n1_untransposed_staged.json 2> gpurun_out/r2_bench_cfg5_n1_untransposed_staged.err
  cut -c1-260 gpurun_out/r2_bench_cfg5_n1_untransposed_staged.json
  cut -c1-260 gpurun_out/r2_bench_cfg5_n1_dgradkn.json gpurun_out/r2_bench_cfg5_n1_dgradkn_staged.json
  # headline kernel: 64-thread CTA variant of the persistent MLP kernel next to the default (128 threads)
  timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/r2_microbench_mlp.json > gpurun_out/r2_microbench_mlp.log 2>&1
  grep variant gpurun_out/r2_microbench_mlp.log | cut -c1-220
  COLEARN_BENCH_E2E_ONE_CALL=1 timeout 60 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg2_n1.json 2> gpurun_out/r2_bench_cfg2_n1.err
  COLEARN_MLP_VARIANT=5 timeout 60 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg2_n1_variant5.json 2> gpurun_out/r2_bench_cfg2_n1_variant5.err
  COLEARN_MLP_VARIANT=4 timeout 60 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg2_n1_variant4.json 2> gpurun_out/r2_bench_cfg2_n1_variant4.err
  cut -c1-200 gpurun_out/r2_bench_cfg2_n1.json gpurun_out/r2_bench_cfg2_n1_variant5.json gpurun_out/r2_bench_cfg2_n1_variant4.json
fi
