mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. encrypted demo on 2 GPUs: dealer + party 0 on GPU 0, party 1 on GPU 1, opens through p2p_copy_kernel
timeout 200 $TR --nproc-per-node 2 --master-port 29831 federated_coordinator.py -t topic/state --box -e --model ffnn --synthetic 256 -w 1 \
    --checkpoint gpurun_out/rA_enc.pth --batch-size 4 --enc-items 64 --log-interval 4 > gpurun_out/rA_box_enc.log 2>&1; echo "box -e rc=$?"
grep -E "share holder|End encryption|Loss:" gpurun_out/rA_box_enc.log | tail -n 5 | cut -c1-220; tail -n 4 gpurun_out/rA_box_enc.log | cut -c1-300
# 2. fused wgrad -> reduce at full size, 2 ranks, short watchdog: which wait fails?
COLEARN_OVERLAP_REDUCE=1 COLEARN_OVERLAP_TIMEOUT_S=6 COLEARN_SPIN_TIMEOUT_S=12 timeout 200 $TR --nproc-per-node 2 --master-port 29832 bench.py --gpus 2 --steps 4 --warmup 3 --config cfg5 --no-e2e \
    > gpurun_out/rA_cfg5_overlap_n2.json 2> gpurun_out/rA_cfg5_overlap_n2.err; echo "overlap n2 rc=$?"
grep -E "colearn:|still waits|twoshot_overlap" gpurun_out/rA_cfg5_overlap_n2.json gpurun_out/rA_cfg5_overlap_n2.err | head -n 6 | cut -c1-300; cut -c1-200 gpurun_out/rA_cfg5_overlap_n2.json | tail -n 2
COLEARN_OVERLAP_REDUCE=1 COLEARN_CUDA_GRAPHS=0 COLEARN_OVERLAP_TIMEOUT_S=6 COLEARN_SPIN_TIMEOUT_S=12 timeout 200 $TR --nproc-per-node 2 --master-port 29833 bench.py --gpus 2 --steps 4 --warmup 3 --config cfg5 --no-e2e \
    > gpurun_out/rA_cfg5_overlap_n2_nographs.json 2> gpurun_out/rA_cfg5_overlap_n2_nographs.err; echo "overlap n2 (no graphs) rc=$?"
grep -E "colearn:|still waits|twoshot_overlap" gpurun_out/rA_cfg5_overlap_n2_nographs.json gpurun_out/rA_cfg5_overlap_n2_nographs.err | head -n 6 | cut -c1-300; cut -c1-200 gpurun_out/rA_cfg5_overlap_n2_nographs.json | tail -n 2
# 3. the default line at N=2 with the per-net gather form
timeout 200 $TR --nproc-per-node 2 --master-port 29834 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/rA_bench_default_n2.json 2> gpurun_out/rA_bench_default_n2.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/rA_bench_default_n2.json') if l.startswith('{')][-1]
print('ref_local n2', d['value'], d['e2e']['value'], d['config']['self_check']['ok'], '| cfg2', d['config']['also_measured']['cfg2']['value'])"
# 4. NVLink counters of the comm kernels
sh scripts/prof_comm_n2.sh 2>&1 | tail -n 14 | cut -c1-250
