#!/bin/sh
# Router-side MUD watcher: keeps the coordinator's IoT allow-list in sync with osMUD.
#
# Runs ON THE ROUTER (OpenWRT + osMUD).  Every INTERVAL seconds it walks the osMUD/dnsmasq event log
# (pipe-separated: field 2 = NEW|DEL|OLD, field 7 = MUD URL or "-", field 10 = device IP) and, for every
# device that presented a MUD URL, asks the coordinator host to add/remove the IP from its allow-list:
#     ssh <user>@<coordinator> "cd <path>; python file_upgrader.py -c NEW|DEL -i <ip>"
# Password-less ssh (dropbear key) from router to coordinator is required.
# Same role as the reference's device_filtering/monitoring_dnsmasq.sh (options -u -p -i -h).

usage() {
    echo "Usage: $0 -u <coordinator user> -p <path of file_upgrader.py on the coordinator> [-i <seconds>] [-c <coordinator host>] [-l <log file>]"
    echo "   -u   coordinator user"
    echo "   -p   directory holding file_upgrader.py on the coordinator"
    echo "   -i   scan interval in seconds (default 10)"
    echo "   -c   coordinator host name (default www.mfs.example.com, as registered in /etc/hosts)"
    echo "   -l   event log to scan (default /var/log/dhcpmasq.txt)"
    echo "   -n   dry run: print the commands instead of running ssh"
    echo "   -1   single pass (no loop)"
    echo "   -h   this help"
}

if [ $# -eq 0 ]; then
    echo "Missing options! (run $0 -h for help)"
    exit 0
fi

REMOTE_USER=""
REMOTE_PATH=""
INTERVAL=10
COORDINATOR="www.mfs.example.com"
LOGFILE="/var/log/dhcpmasq.txt"
DRY=0
ONCE=0
while getopts "hu:p:i:c:l:n1" OPTION; do
    case $OPTION in
        u) REMOTE_USER=$OPTARG ;;
        p) REMOTE_PATH=$OPTARG ;;
        i) INTERVAL=$OPTARG ;;
        c) COORDINATOR=$OPTARG ;;
        l) LOGFILE=$OPTARG ;;
        n) DRY=1 ;;
        1) ONCE=1 ;;
        h) usage; exit 0 ;;
        *) usage; exit 1 ;;
    esac
done

scan() {
    [ -r "$LOGFILE" ] || { echo "cannot read $LOGFILE"; return; }
    while IFS= read -r line; do
        COMMAND=$(echo "$line" | awk -F "|" '{ print $2 }')
        MUD_URL=$(echo "$line" | awk -F "|" '{ print $7 }')
        IP=$(echo "$line" | awk -F "|" '{ print $10 }')
        if [ "$MUD_URL" = "-" ] || [ "$COMMAND" = "OLD" ] || [ -z "$IP" ]; then
            echo "not valid: $line"
            continue
        fi
        CMD="cd $REMOTE_PATH/; python file_upgrader.py -c $COMMAND -i $IP"
        if [ "$DRY" -eq 1 ]; then
            echo "ssh $REMOTE_USER@$COORDINATOR \"$CMD\""
        else
            ssh "$REMOTE_USER@$COORDINATOR" "$CMD" < /dev/null
        fi
    done < "$LOGFILE"
}

while true; do
    scan
    [ "$ONCE" -eq 1 ] && break
    sleep "$INTERVAL"
done
