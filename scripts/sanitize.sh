#!/bin/sh
# compute-sanitizer passes over the single-GPU kernel tests (SURVEY §5 "Race detection / sanitizers").
# Usage (under gpurun): sh scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]   -> gpurun_out/sanitizer_<tool>_*.log
# Three targets per tool: the persistent MLP + elementwise kernels, the tcgen05 GEMM family (plain / MN-major / split-K), and
# the conv / BatchNorm / pooling kernels.  (The star / two-shot kernels need peers: tests/test_multigpu.py under memcheck.)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
run() {  # name, pytest args...
  name=$1; shift
  # (--report-api-errors no: the lazy kernel lookup inside cudart returns CUDA_ERROR_INVALID_HANDLE once per kernel and retries; not an error of ours)
  timeout 600 compute-sanitizer --tool "$TOOL" --report-api-errors no --error-exitcode 1 --log-file "gpurun_out/sanitizer_${TOOL}_${name}.log" \
      python -m pytest "$@" -m gpu -q -x --timeout 900 -p no:cacheprovider > "gpurun_out/sanitizer_${TOOL}_${name}.out" 2>&1
  echo "rc=$?" >> "gpurun_out/sanitizer_${TOOL}_${name}.log"
  echo "== $TOOL $name: $(tail -n 1 gpurun_out/sanitizer_${TOOL}_${name}.log) | $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_${TOOL}_${name}.log) summary line(s): $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_${TOOL}_${name}.log | tail -n 1)"
  tail -n 2 "gpurun_out/sanitizer_${TOOL}_${name}.out"
}
run mlp tests/test_gpu_kernels.py -k "persistent_mlp_matches_reference and 97 or in_kernel_shuffle or losses or sgd_and_fedavg or eval_argmax"
run gemm tests/test_gpu_kernels.py tests/test_gpu_schedules.py -k "gemm_tcgen05_plain and 256 or gemm_mn_major_fused or gemm_split_k_partials and 4096 or splitk_reduce_kernel"
run conv tests/test_conv_ops.py
