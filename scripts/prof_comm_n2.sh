#!/bin/sh
# NVLink evidence for the cross-GPU kernels (2 GPUs of one box): rank 1 runs plainly, rank 0 runs under ncu with a handful of
# counters (kernel time, NVLink bytes sent / received by this GPU, L2 / DRAM bytes) on star_round_kernel and
# twoshot_fedavg_kernel.  ncu replays a kernel once per counter group; the kernels tolerate that (flags are monotonic epochs, a
# replay finds them already raised; the two-shot reduce re-reads peer data that its peer no longer changes).
#   -> gpurun_out/r2_prof_twoshot_n2.ncu-rep, gpurun_out/r2_prof_star_n2.ncu-rep  (+ .csv exports)
mkdir -p gpurun_out
M="gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,lts__t_bytes.sum,dram__bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed"
export MASTER_ADDR=127.0.0.1 WORLD_SIZE=2 COLEARN_SPIN_TIMEOUT_S=300
run_pair() {  # name, kernel regex, count, command...
  name=$1; regex=$2; count=$3; shift 3
  export MASTER_PORT=$((29960 + $(echo "$name" | wc -c)))
  ( RANK=1 LOCAL_RANK=1 timeout 300 "$@" > gpurun_out/r2_prof_${name}_rank1.log 2>&1 ) &
  RANK=0 LOCAL_RANK=0 timeout 300 ncu --metrics "$M" --clock-control none -k "regex:$regex" -s 4 -c "$count" -f -o gpurun_out/r2_prof_${name}_n2 \
      "$@" > gpurun_out/r2_prof_${name}_rank0.log 2>&1
  echo "$name ncu rc=$?"
  wait
  ncu -i gpurun_out/r2_prof_${name}_n2.ncu-rep --page raw --csv > gpurun_out/r2_prof_${name}_n2.csv 2>/dev/null
  python - "$name" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r2_prof_{name}_n2.csv")))
if len(rows) > 2:
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        t = float(d.get("gpu__time_duration.sum", "0").replace(",", "")) or 1.0
        unit = rows[1][h.index("gpu__time_duration.sum")]
        tx, rx = float(d.get("nvltx__bytes.sum", "0").replace(",", "")), float(d.get("nvlrx__bytes.sum", "0").replace(",", ""))
        print(name, d.get("Kernel Name", "")[:40], "time", t, unit, "nvl tx", tx, rows[1][h.index("nvltx__bytes.sum")], "rx", rx)
PY
}
run_pair twoshot "twoshot_fedavg_kernel" 3 python scripts/comm_sweep.py --sizes 50397188 --out gpurun_out/r2_prof_twoshot_sweep.json
run_pair twoshot_p2p "twoshot_fedavg_kernel" 3 python scripts/comm_sweep.py --sizes 50397188 --nvls 0 --out gpurun_out/r2_prof_twoshot_p2p_sweep.json
run_pair star "star_round_kernel" 6 python bench.py --gpus 2 --steps 6 --warmup 3 --config cfg2 --no-e2e
