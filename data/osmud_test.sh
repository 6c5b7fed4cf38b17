#!/bin/sh
# osMUD start-up performance probe for an OpenWRT router (same role as the reference's data/osmud_test.sh):
#   $1 = repetitions, $2 = seconds to let osMUD run each time.  Collects /var/log/osmud_perf.log per run.
N=${1:-1}
WAIT=${2:-30}
mkdir -p result
i=0
while [ "$i" -lt "$N" ]; do
    cp /rom/etc/config/firewall /etc/config/firewall && /etc/init.d/firewall restart
    echo "firewall restarted; starting osmud"
    ./startup.sh &
    sleep "$WAIT"
    /etc/init.d/osmud stop
    i=$((i + 1))
    cp /var/log/osmud_perf.log "result/test_$i.txt"
done
