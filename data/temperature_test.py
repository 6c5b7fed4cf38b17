#!/usr/bin/env python
"""Temperature monitor kept at the reference's path (``data/temperature_test.py``): samples
``psutil.sensors_temperatures()`` once per second into ``monitoring_temp.txt`` until Ctrl-C."""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.utils.monitors import monitor_temperature  # noqa: E402

if __name__ == "__main__":
    stop = threading.Event()
    print("Monitoring temperature started")
    try:
        monitor_temperature("monitoring_temp.txt", 1.0, stop)
    except KeyboardInterrupt:
        stop.set()
        print("Monitoring temperature ended")
