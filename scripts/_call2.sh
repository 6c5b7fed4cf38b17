mkdir -p gpurun_out
timeout 100 python scripts/debug_splitk_step.py COLEARN_CONV_SPLITK=1 > gpurun_out/r2_debug_splitk.json 2> gpurun_out/r2_debug_splitk.err; echo "splitk dbg rc=$?"
head -c 2500 gpurun_out/r2_debug_splitk.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r2_bench_reference_n1.json; tail -n 5 gpurun_out/r2_bench_reference_n1.err
sh scripts/run_round2_first.sh gemm
sh scripts/run_round2_first.sh conv
