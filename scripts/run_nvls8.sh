#!/bin/sh
# 8-GPU comm experiment: two-shot FedAvg kernel with NVLS (multimem.ld_reduce/st) vs P2P, CTA counts, + cfg5 bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
SZ="1048576,11181644,50397188"
timeout 150 $TR --master-port 29701 scripts/comm_sweep.py --shadow --nvls 1 --sizes $SZ --tag nvls --out gpurun_out/n8_nvls.json > gpurun_out/n8_nvls.log 2>&1
timeout 150 $TR --master-port 29702 scripts/comm_sweep.py --shadow --nvls 0 --sizes $SZ --tag p2p --out gpurun_out/n8_p2p.json > gpurun_out/n8_p2p.log 2>&1
timeout 150 $TR --master-port 29703 scripts/comm_sweep.py --shadow --nvls 1 --blocks 148 --sizes $SZ --tag nvls148 --out gpurun_out/n8_nvls148.json > gpurun_out/n8_nvls148.log 2>&1
timeout 150 $TR --master-port 29704 scripts/comm_sweep.py --nvls 1 --sizes $SZ --tag nvls_noshadow --out gpurun_out/n8_nvls_ns.json > gpurun_out/n8_nvls_ns.log 2>&1
timeout 200 $TR --master-port 29705 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > gpurun_out/n8_cfg5.json 2> gpurun_out/n8_cfg5.err
COLEARN_NVLS=0 timeout 200 $TR --master-port 29706 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > gpurun_out/n8_cfg5_p2p.json 2> gpurun_out/n8_cfg5_p2p.err
grep -h twoshot_ms gpurun_out/n8_*.log | cut -c1-400
cut -c1-220 gpurun_out/n8_cfg5.json gpurun_out/n8_cfg5_p2p.json
tail -n 3 gpurun_out/n8_cfg5.err | cut -c1-300
