#!/bin/sh
# The reference ships the raw logs of its published experiments (data/{1000,2000,3000}/*round*batches/coordinator/round_testing*.txt,
# data/README.md).  This script re-runs those experiments with THIS repo's CLIs on the box it is called on — real coordinator,
# embedded bus broker, two remote_worker.py processes over TCP, R rounds x N local batch-1 iterations, BCE, lr 0.01 — and leaves
# logs in the same format under OUT (default gpurun_out/experiment_logs):  <N>/<R>round<N>batches/coordinator/round_testing_<rep>.txt
# plus the coordinator's CPU monitor, like the reference's folder layout.
OUT=${1:-gpurun_out/experiment_logs}
REPS=${2:-3}
PORT=18960
NS=${NS:-"1000 2000 3000"}
RS=${RS:-"3 6 12"}
for N in $NS; do
  for R in $RS; do
    [ "$N" != 1000 ] && [ "$R" = 12 ] && continue        # the reference measured 12 rounds only with 1000 iterations
    D="$OUT/$N/${R}round${N}batches/coordinator"
    mkdir -p "$D"
    rep=1
    while [ $rep -le "$REPS" ]; do
      PORT=$((PORT + 3))
      python federated_coordinator.py -t topic/state -w 1 -r -f "$R" --max-batches "$N" --embedded-broker -p "$PORT" --host 127.0.0.1 \
          --exit-after 1 --checkpoint "$D/test_$rep.pth" --round-log "$D/round_testing_$rep.txt" --evaluate --synthetic 4096 \
          > "$D/coordinator_$rep.log" 2>&1 &
      CPID=$!
      sleep 2
      python data/ps_util_test.py -p "$CPID" > /dev/null 2>&1 &
      MPID=$!
      python remote_worker.py --host 127.0.0.1 -p $((PORT + 1)) -b 127.0.0.1 --broker-port "$PORT" -t topic/state -w 1 --synthetic 4096 --seed 1 > "$D/pi1_$rep.log" 2>&1 &
      W1=$!
      python remote_worker.py --host 127.0.0.1 -p $((PORT + 2)) -b 127.0.0.1 --broker-port "$PORT" -t topic/state -w 1 --synthetic 4096 --seed 2 > "$D/pi2_$rep.log" 2>&1 &
      W2=$!
      ( sleep 120; kill $CPID 2>/dev/null ) &
      WD=$!
      wait $CPID
      kill $W1 $W2 $MPID $WD 2>/dev/null
      [ -f monitoring_cpu.txt ] && mv monitoring_cpu.txt "$D/monitoring_cpu_$rep.txt"
      rm -f "$D/test_$rep.pth" "$D/test_$rep.pth.json"
      tail -n 2 "$D/round_testing_$rep.txt" 2>/dev/null | tr '\n' ' '; echo " <- $D/round_testing_$rep.txt"
      rep=$((rep + 1))
    done
  done
done
