#!/bin/sh
# osMUD start-up performance probe for an OpenWRT router.
#
# Each repetition resets the firewall to the factory rules, lets osMUD run for a fixed time, stops it and keeps
# that run's performance log.  Same purpose as the reference's data/osmud_test.sh (positional: repetitions, seconds),
# with named options, a dry-run mode for boxes that are not routers, and a summary line per run.
#
#   osmud_test.sh [-n RUNS] [-t SECONDS] [-o OUTDIR] [-s STARTUP_SCRIPT] [-d]      (or: osmud_test.sh RUNS SECONDS)
runs=1
seconds=30
outdir=result
startup=./startup.sh
dry=0
perf_log=/var/log/osmud_perf.log

usage() { sed -n '2,9p' "$0" | sed 's/^# \{0,1\}//'; exit "${1:-0}"; }

while getopts "n:t:o:s:dh" opt; do
    case "$opt" in
        n) runs=$OPTARG ;;
        t) seconds=$OPTARG ;;
        o) outdir=$OPTARG ;;
        s) startup=$OPTARG ;;
        d) dry=1 ;;
        h) usage 0 ;;
        *) usage 2 ;;
    esac
done
shift $((OPTIND - 1))
[ -n "$1" ] && runs=$1
[ -n "$2" ] && seconds=$2

run() {    # echo the command in dry-run mode, execute it otherwise
    if [ "$dry" -eq 1 ]; then echo "+ $*"; else "$@"; fi
}

reset_firewall() {
    run cp /rom/etc/config/firewall /etc/config/firewall
    run /etc/init.d/firewall restart
}

one_run() {
    idx=$1
    reset_firewall
    if [ "$dry" -eq 1 ]; then echo "+ $startup &"; else "$startup" & fi
    run sleep "$seconds"
    run /etc/init.d/osmud stop
    dest="$outdir/test_$idx.txt"
    run cp "$perf_log" "$dest"
    echo "run $idx/$runs: osmud ran ${seconds}s, log kept in $dest"
}

[ "$dry" -eq 1 ] || mkdir -p "$outdir"
n=1
while [ "$n" -le "$runs" ]; do
    one_run "$n"
    n=$((n + 1))
done
