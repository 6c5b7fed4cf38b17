#!/usr/bin/env python
"""Thin launcher kept at the reference's path (``device_filtering/file_upgrader.py``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.tools.file_upgrader import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
