#!/bin/sh
# One-box scaling sweep (run under `gpurun --gpus 8`): multi-rank tests, bench at N=1/2/4/8 for the flagship
# config, the other BASELINE configs at N=8, the torch+NCCL comparator and the comm bandwidth sweep.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 280 -k multi_rank > gpurun_out/s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s_tests.log
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/s_cfg2_n1.json 2> gpurun_out/s_cfg2_n1.err
P=29600
for N in 2 4 8; do
  P=$((P+1))
  timeout 200 $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s_cfg2_n$N.json 2> gpurun_out/s_cfg2_n$N.err
done
timeout 200 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 --impl torch_nccl > gpurun_out/s_cfg2_n8_nccl.json 2> gpurun_out/s_cfg2_n8_nccl.err
timeout 200 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg3 > gpurun_out/s_cfg3_n8.json 2> gpurun_out/s_cfg3_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > gpurun_out/s_cfg5_n8.json 2> gpurun_out/s_cfg5_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29614 bench.py --gpus 8 --steps 5 --warmup 3 --config cfg5 --impl torch_nccl > gpurun_out/s_cfg5_n8_nccl.json 2> gpurun_out/s_cfg5_n8_nccl.err
timeout 200 $TR --nproc-per-node 8 --master-port 29615 bench.py --gpus 8 --steps 5 --warmup 3 --config cfg4 > gpurun_out/s_cfg4_n8.json 2> gpurun_out/s_cfg4_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29616 scripts/comm_sweep.py --shadow --out gpurun_out/s_sweep_n8.json > gpurun_out/s_sweep_n8.log 2>&1
timeout 200 $TR --nproc-per-node 4 --master-port 29617 scripts/comm_sweep.py --shadow --out gpurun_out/s_sweep_n4.json > gpurun_out/s_sweep_n4.log 2>&1
timeout 200 $TR --nproc-per-node 8 --master-port 29618 scripts/comm_sweep.py --out gpurun_out/s_sweep_n8_noshadow.json > gpurun_out/s_sweep_n8_noshadow.log 2>&1
tail -n 6 gpurun_out/s_tests.log
for f in gpurun_out/s_cfg*.json; do echo "== $f"; cut -c1-260 $f; done
grep -h twoshot_ms gpurun_out/s_sweep_n8.log | cut -c1-330
tail -n 3 gpurun_out/s_cfg2_n8.err gpurun_out/s_cfg5_n8.err gpurun_out/s_cfg4_n8.err | cut -c1-300
