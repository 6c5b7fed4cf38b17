#!/bin/sh
# One 8-GPU box (gpurun --gpus 8): the multi-rank GPU tests at world 8, the default bench line (both arms' config) and
# BASELINE configs 2-5 at N = 4 and 8 (N = 1, 2 come from the cheaper 1- / 2-GPU calls), the torch+NCCL comparator, the
# broadcast + FedAvg-reduce bandwidth sweep next to dist.all_reduce, and box mode as a service on 8 GPUs.
# Independent jobs of <= 4 ranks run side by side on disjoint halves of the box.  Everything lands in gpurun_out/r2_8_*.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out/r2_8
( COLEARN_OVERLAP_TIMEOUT_S=8 timeout 420 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 200 -p no:cacheprovider -k "multi_rank and (star or nvls_weighted or wide or deadline)" > ${O}_tests.log 2>&1; echo "rc=$?" >> ${O}_tests.log ) 
tail -n 6 ${O}_tests.log | cut -c1-300
# --- N = 8 ---------------------------------------------------------------------------------------------------------
timeout 200 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 3 > ${O}_default_n8.json 2> ${O}_default_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg3 > ${O}_cfg3_n8.json 2> ${O}_cfg3_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 5 --warmup 3 --config cfg4 > ${O}_cfg4_n8.json 2> ${O}_cfg4_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > ${O}_cfg5_n8.json 2> ${O}_cfg5_n8.err
COLEARN_OVERLAP_REDUCE=1 timeout 200 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg5 > ${O}_cfg5_n8_overlap16.json 2> ${O}_cfg5_n8_overlap16.err
timeout 200 $TR --nproc-per-node 8 --master-port 29709 scripts/comm_sweep.py --out ${O}_sweep_n8_nvls.json > ${O}_sweep_n8_nvls.log 2>&1
timeout 200 $TR --nproc-per-node 8 --master-port 29710 scripts/comm_sweep.py --nvls 0 --out ${O}_sweep_n8_p2p.json > ${O}_sweep_n8_p2p.log 2>&1
timeout 120 $TR --nproc-per-node 8 --master-port 29711 federated_coordinator.py -t topic/state --box --model mlp --synthetic 8192 -w 1 --checkpoint ${O}_box.pth \
    --exit-after 2 --select 4 --selection first -f 10 --local-epochs 5 --metrics ${O}_box_cfg3_rounds.jsonl > ${O}_box_cfg3.log 2>&1
echo "box rc=$?"; grep -E "window closed|Total training" ${O}_box_cfg3.log | cut -c1-260
# --- N = 4 (two jobs side by side on GPUs 0-3 / 4-7) -----------------------------------------------------------------
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 $TR --nproc-per-node 4 --master-port 29721 bench.py --gpus 4 --steps 20 --warmup 3 > ${O}_default_n4.json 2> ${O}_default_n4.err ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 200 $TR --nproc-per-node 4 --master-port 29722 bench.py --gpus 4 --steps 5 --warmup 3 --config cfg4 > ${O}_cfg4_n4.json 2> ${O}_cfg4_n4.err ) &
wait
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 $TR --nproc-per-node 4 --master-port 29723 bench.py --gpus 4 --steps 10 --warmup 3 --config cfg5 > ${O}_cfg5_n4.json 2> ${O}_cfg5_n4.err ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 200 $TR --nproc-per-node 2 --master-port 29724 bench.py --gpus 2 --steps 5 --warmup 3 --config cfg4 > ${O}_cfg4_n2.json 2> ${O}_cfg4_n2.err ) &
( CUDA_VISIBLE_DEVICES=6,7 timeout 200 $TR --nproc-per-node 2 --master-port 29725 scripts/comm_sweep.py --out ${O}_sweep_n2_nvls.json > ${O}_sweep_n2_nvls.log 2>&1 ) &
wait
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_8_*.json")):
    try:
        lines = [l for l in open(f).read().splitlines() if l.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    if "value" in d:
        c = d.get("config", {})
        print(f.split("r2_8_")[1], "N=%s value=%.2f e2e=%s ms=%.3f path=%s nvls=%s check=%s roofline=%s" % (
            d.get("n_gpus"), d["value"], (d.get("e2e") or {}).get("value"), d["ms_per_step"], c.get("train_path"), c.get("nvls"),
            (c.get("self_check") or {}).get("ok"), (c.get("roofline") or {}).get("roofline_frac")),
            "| cfg2:", (c.get("also_measured", {}).get("cfg2") or {}).get("value"))
for f in sorted(glob.glob("gpurun_out/r2_8_sweep_*.json")):
    for r in json.load(open(f)):
        print(f.split("r2_8_")[1], r["P"], "ours %.1f GB/s (%.3f of 900) nccl %.1f nvls=%s" % (r["twoshot_busbw_GBps"] or 0, r["twoshot_frac_of_900GBps"] or 0, r["nccl_busbw_GBps"] or 0, r["nvls"]))
PY
tail -n 4 ${O}_default_n8.err ${O}_cfg5_n8_overlap16.err ${O}_cfg4_n8.err | cut -c1-300
