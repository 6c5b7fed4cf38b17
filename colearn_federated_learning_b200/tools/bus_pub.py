"""``mosquitto_pub`` stand-in: ``python -m colearn_federated_learning_b200.tools.bus_pub -t topic/state
-m "(192.168.1.7, TRAINING)" [-h host] [-p port]`` (reference usage: README.md:80, fc.py:3)."""
import argparse
import time

from ..control.bus import BusClient


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--help", action="help")
    ap.add_argument("-t", "--topic", required=True)
    ap.add_argument("-m", "--message", required=True)
    ap.add_argument("-h", "--host", default="localhost")
    ap.add_argument("-p", "--port", type=int, default=1883)
    ns = ap.parse_args(argv)
    c = BusClient("bus_pub", transport="tcp")
    c.connect(ns.host, ns.port)
    c.publish(ns.topic, ns.message)
    time.sleep(0.05)  # let the frame leave before the socket closes
    c.disconnect()


if __name__ == "__main__":
    main()
