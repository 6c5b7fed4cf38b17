mkdir -p gpurun_out
timeout 60 python scripts/debug_overlap.py 512 3 128 > gpurun_out/r2c6_debug_overlap_small.json 2> gpurun_out/r2c6_debug_overlap_small.err; echo "overlap small rc=$?"; cut -c1-1500 gpurun_out/r2c6_debug_overlap_small.json; tail -n 3 gpurun_out/r2c6_debug_overlap_small.err | cut -c1-300
timeout 60 python scripts/debug_overlap.py 4096 4 1024 > gpurun_out/r2c6_debug_overlap_wide.json 2> gpurun_out/r2c6_debug_overlap_wide.err; echo "overlap wide rc=$?"; cut -c1-1500 gpurun_out/r2c6_debug_overlap_wide.json
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_schedules.py::test_overlapped_reduce_single_gpu_is_bit_identical > gpurun_out/r2c6_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2c6_pytest_gpu.log | cut -c1-220
timeout 120 python scripts/microbench.py --only mlp --out gpurun_out/r2c6_microbench_mlp.json > gpurun_out/r2c6_microbench_mlp.log 2>&1; grep "'batch': 1, 'samples': 8192" gpurun_out/r2c6_microbench_mlp.log | cut -c40-175
REPS=1; sh scripts/paper_run_logs.sh gpurun_out/experiment_logs 1 2>&1 | tail -n 8
sh scripts/sanitize.sh memcheck 2>&1 | grep "=="
