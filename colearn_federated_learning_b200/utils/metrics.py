"""Round metrics: structured JSONL + the reference's human-readable round log.

The reference's experiment logs (``data/*/coordinator/round_testing*.txt``) have the shape
``Testing round on: <max_batches>`` / ``Loss for worker id: <id> tensor(...)`` /
``Time round <r> : <sec>`` / ``Total training time: <sec>`` /
``Loss evaluation global model after training: <x>`` — but the code that wrote
them is not in the reference tree.  :class:`RoundLogger` writes exactly that
format plus one JSON object per round for machines.
"""
from __future__ import annotations

import json
import logging
import os
import threading
import time
from typing import Any, Dict, List, Optional

log = logging.getLogger(__name__)


class RoundLogger:
    def __init__(self, jsonl_path: Optional[str] = None, text_path: Optional[str] = None) -> None:
        self.jsonl_path, self.text_path = jsonl_path, text_path
        self._lock = threading.Lock()
        self.records: List[Dict[str, Any]] = []
        self._t_start: Optional[float] = None
        for p in (jsonl_path, text_path):
            if p:
                os.makedirs(os.path.dirname(os.path.abspath(p)), exist_ok=True)

    def _text(self, line: str) -> None:
        log.info(line)
        if self.text_path:
            with open(self.text_path, "a") as f:
                f.write(line + "\n")

    def start_training(self, max_batches: int) -> None:
        self._t_start = time.time()
        self._text(f"Testing round on: {max_batches}")

    def worker_loss(self, worker_id: str, loss: float) -> None:
        self._text(f"Loss for worker id: {worker_id} tensor({loss:.4f}, requires_grad=True)")

    def end_round(self, round_idx: int, seconds: float, **fields: Any) -> Dict[str, Any]:
        self._text(f"Time round {round_idx} : {seconds}")
        rec = {"round": round_idx, "t_round_s": seconds, "ts": time.time(), **fields}
        with self._lock:
            self.records.append(rec)
            if self.jsonl_path:
                with open(self.jsonl_path, "a") as f:
                    f.write(json.dumps(rec, default=float) + "\n")
        return rec

    def end_training(self, eval_loss: Optional[float] = None) -> float:
        total = time.time() - (self._t_start or time.time())
        self._text(f"Total training time: {total}")
        if eval_loss is not None:
            self._text(f"Loss evaluation global model after training: {eval_loss}")
        return total
