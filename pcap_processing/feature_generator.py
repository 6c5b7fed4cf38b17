#!/usr/bin/env python
"""Launcher kept at the reference's path (``pcap_processing/feature_generator.py``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.tools.feature_generator import main  # noqa: E402

if __name__ == "__main__":
    main()
