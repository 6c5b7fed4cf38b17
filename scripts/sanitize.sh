#!/bin/sh
# compute-sanitizer passes over the single-GPU kernel tests (SURVEY §5 "Race detection / sanitizers").
# Usage (under gpurun): sh scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
TOOL=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --error-exitcode 1 --log-file "gpurun_out/sanitizer_$TOOL.log" \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "persistent_mlp_matches_reference and 97 or losses or sgd_and_fedavg or eval_argmax" --timeout 900
echo "rc=$?" >> "gpurun_out/sanitizer_$TOOL.log"
