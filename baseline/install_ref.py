"""Put the UNMODIFIED reference under ``baseline/_ref`` (git-ignored; it travels to the GPU box with the snapshot).

``pip install --target baseline/_ref /root/reference`` cannot work — the reference is a flat tree of scripts without
``setup.py`` / ``pyproject.toml`` (pip: "Directory '/root/reference' is not installable", also from a /tmp copy and with
``--no-deps``).  Its modules are meant to be run from the checkout, so the "install" is a byte-identical copy of the
Python sources + the example dataset + the allow-list, with a ``SHA256SUMS`` manifest that ``verify()`` re-checks before
every reference-arm run.  Nothing here imports the product package.

    python baseline/install_ref.py [--src /root/reference]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
from typing import Dict

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["federated_coordinator.py", "client_federated.py", "remote_worker.py", "datasets.py", "event_parser.py", "settings.py",
         "README.md", "device_filtering/filtering_file.txt", "device_filtering/file_upgrader.py",
         "dataset_example/UNSW_2018_IoT_Botnet_Final_10_best_Training_1_1.csv"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src: str = "/root/reference", dest: str = DEST) -> Dict[str, str]:
    sums = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        sums[rel] = _sha(d)
        assert sums[rel] == _sha(s)
    with open(os.path.join(dest, "SHA256SUMS"), "w") as f:
        for rel, h in sorted(sums.items()):
            f.write(f"{h}  {rel}\n")
    return sums


def verify(dest: str = DEST) -> Dict[str, str]:
    """Re-hash the installed files against the manifest written at install time; returns {file: sha256}."""
    manifest = os.path.join(dest, "SHA256SUMS")
    if not os.path.exists(manifest):
        raise FileNotFoundError(f"{manifest} missing: run `python baseline/install_ref.py` where /root/reference is mounted")
    sums = {}
    with open(manifest) as f:
        for line in f:
            h, rel = line.strip().split("  ", 1)
            if _sha(os.path.join(dest, rel)) != h:
                raise RuntimeError(f"baseline/_ref/{rel} differs from the installed reference file")
            sums[rel] = h
    return sums


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    for rel, h in sorted(install(src).items()):
        print(h, rel)
