"""IoT allow-list updater (reference ``device_filtering/file_upgrader.py:5-45``).

``-c NEW -i <ip>`` appends the address if absent; ``-c DEL -i <ip>`` removes it.  The router-side
MUD watcher (``monitoring_dnsmasq.sh``) calls this over ssh.  The reference's DEL path writes
through the wrong file handle and duplicates instead of deleting; here DEL
rewrites the file atomically.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
from typing import List

DEFAULT_FILE = "filtering_file.txt"


def read_ips(path: str) -> List[str]:
    try:
        with open(path) as f:
            return [line.rstrip() for line in f if line.rstrip()]
    except FileNotFoundError:
        return []


def _write(path: str, ips: List[str]) -> None:
    d = os.path.dirname(os.path.abspath(path))
    fd, tmp = tempfile.mkstemp(dir=d, prefix=".filter-")
    with os.fdopen(fd, "w") as f:
        for ip in ips:
            f.write(ip + "\n")
    os.replace(tmp, path)


def add_ip(ip: str, path: str) -> bool:
    ips = read_ips(path)
    if ip in ips:
        print("This ip already exist")
        return False
    _write(path, ips + [ip])
    print("done")
    return True


def delete_ip(ip: str, path: str) -> bool:
    ips = read_ips(path)
    if ip not in ips:
        return False
    print("Deleting ip address")
    ips.remove(ip)
    _write(path, ips)
    return True


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="File upgrader execution: the output file is 'filtering_file.txt'")
    parser.add_argument("--command", "-c", type=str, default="NEW", help="Insert or delete an ip address from the filtering file")
    parser.add_argument("--ip", "-i", type=str, required=True, help="Ip address to add or remove from the filtering_file.txt")
    parser.add_argument("--file", "-f", type=str, default=None, help="allow-list path (default: ./filtering_file.txt)")
    return parser


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    path = args.file or os.path.join(os.getcwd(), DEFAULT_FILE)
    if args.command == "NEW":
        print("Trying to insert new ip address")
        add_ip(args.ip, path)
    elif args.command == "DEL":
        delete_ip(args.ip, path)
    else:
        print("Not recognized command")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
