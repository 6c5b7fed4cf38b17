#!/bin/sh
# Round-2 one-box measurement (gpurun --gpus 8), AFTER scripts/run_round2_first.sh and the SCHEDULE_DEFAULTS flip:
#   * cfg4 (ResNet-18 bf16, own conv path + fused two-shot FedAvg) at N = 1, 2, 4, 8 — only N=1 was measured in round 1
#   * its torch+NCCL(+cuDNN autocast) comparator at N = 8
#   * the headline (cfg2) and cfg5 at N = 8 again (regression check of the GEMM changes: split-K / MN-major / watchdog)
#   * the multi-rank GPU tests
# Extra schedule flags can be passed through the environment, e.g.  COLEARN_CONV_IMPLICIT=2 sh scripts/run_round2_8gpu.sh
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 280 -k multi_rank > gpurun_out/r2_8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_8_tests.log
tail -n 4 gpurun_out/r2_8_tests.log
timeout 120 python bench.py --gpus 1 --steps 5 --warmup 3 --config cfg4 > gpurun_out/r2_8_cfg4_n1.json 2> gpurun_out/r2_8_cfg4_n1.err
P=29700
for N in 2 4 8; do
  P=$((P+1))
  timeout 200 $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --config cfg4 \
      > gpurun_out/r2_8_cfg4_n$N.json 2> gpurun_out/r2_8_cfg4_n$N.err
done
timeout 200 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 3 --warmup 3 --config cfg4 --impl torch_nccl \
    > gpurun_out/r2_8_cfg4_n8_nccl.json 2> gpurun_out/r2_8_cfg4_n8_nccl.err
timeout 200 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_8_cfg2_n8.json 2> gpurun_out/r2_8_cfg2_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29713 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > gpurun_out/r2_8_cfg5_n8.json 2> gpurun_out/r2_8_cfg5_n8.err
# fused wgrad GEMM -> FedAvg reduce (opt-in): the same cfg5 round with the two-shot kernel running next to the last backward
COLEARN_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 280 -k wide_overlap > gpurun_out/r2_8_overlap_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_8_overlap_test.log
tail -n 3 gpurun_out/r2_8_overlap_test.log
for C in 8 16 32; do
  COLEARN_OVERLAP_REDUCE=1 COLEARN_OVERLAP_CTAS=$C timeout 200 $TR --nproc-per-node 8 --master-port $((29720+C)) bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 \
      > gpurun_out/r2_8_cfg5_n8_overlap$C.json 2> gpurun_out/r2_8_cfg5_n8_overlap$C.err
done
# headline e2e through ONE run_rounds(K) call with the lagged loss read-back (opt-in measurement)
COLEARN_BENCH_E2E_ONE_CALL=1 timeout 200 $TR --nproc-per-node 8 --master-port 29761 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_8_cfg2_n8_e2e_one_call.json 2> gpurun_out/r2_8_cfg2_n8_e2e_one_call.err
for f in gpurun_out/r2_8_*.json; do echo "== $f"; cut -c1-260 $f; done
tail -n 3 gpurun_out/r2_8_cfg4_n8.err | cut -c1-300
