#!/bin/sh
# Final one-box measurement (gpurun --gpus 8): flagship scaling points + wide-MLP config + comm sweep.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/f_cfg2_n1.json 2> gpurun_out/f_cfg2_n1.err
timeout 150 $TR --nproc-per-node 2 --master-port 29801 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/f_cfg2_n2.json 2> gpurun_out/f_cfg2_n2.err
timeout 150 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/f_cfg2_n4.json 2> gpurun_out/f_cfg2_n4.err
timeout 150 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/f_cfg2_n8.json 2> gpurun_out/f_cfg2_n8.err
timeout 150 $TR --nproc-per-node 8 --master-port 29804 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > gpurun_out/f_cfg5_n8.json 2> gpurun_out/f_cfg5_n8.err
timeout 150 $TR --nproc-per-node 8 --master-port 29805 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg3 > gpurun_out/f_cfg3_n8.json 2> gpurun_out/f_cfg3_n8.err
timeout 150 $TR --nproc-per-node 8 --master-port 29806 bench.py --gpus 8 --steps 10 --warmup 3 --config ffnn > gpurun_out/f_ffnn_n8.json 2> gpurun_out/f_ffnn_n8.err
for f in gpurun_out/f_*.json; do echo "== $f"; cut -c1-240 $f; done
tail -n 3 gpurun_out/f_cfg2_n8.err | cut -c1-300
