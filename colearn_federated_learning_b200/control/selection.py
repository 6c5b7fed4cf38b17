"""Worker selection (bounds + policy).

Parity: reference gates training on ``len(devices) >= lower_bound`` (fc.py:320,397,489) and, at
``>= upper_bound``, "applies selection criteria" — a TODO that only logs for local/remote
(fc.py:322-323,491-493) and picks the first two devices in encrypted mode (fc.py:401-404).
Here the policy is real: ``first`` (registration order — what encrypted mode does), ``random``
(seeded), or ``all`` (reference local/remote behaviour), optionally capped by ``--select k``
(BASELINE config 3: "temporal-window selects 4 of 8 workers").
"""
from __future__ import annotations

import logging
import random
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, Optional

log = logging.getLogger(__name__)

LOWER_BOUND = 1        # fc.py:113
LOWER_BOUND_ENC = 2    # fc.py:114
UPPER_BOUND = 100      # fc.py:115
UPPER_BOUND_ENC = 2    # fc.py:116


@dataclass
class SelectionPolicy:
    lower_bound: int = LOWER_BOUND
    upper_bound: int = UPPER_BOUND
    policy: str = "all"            # all | first | random
    select_k: Optional[int] = None  # hard cap independent of upper_bound
    seed: int = 1

    def admits(self, n_devices: int) -> bool:
        return n_devices >= self.lower_bound

    def select(self, devices: Dict[str, Any], round_idx: int = 0) -> "OrderedDict[str, Any]":
        """Return the ordered subset that trains this window; empty if below the lower bound."""
        if not self.admits(len(devices)):
            return OrderedDict()
        items = list(devices.items())
        cap = None
        if len(items) >= self.upper_bound:
            log.info("Applying selection criteria")
            if self.policy != "all":
                cap = self.upper_bound
        if self.select_k is not None:
            cap = self.select_k if cap is None else min(cap, self.select_k)
        if cap is None or cap >= len(items):
            return OrderedDict(items)
        if self.policy == "random":
            rng = random.Random(self.seed * 1000003 + round_idx)
            idx = sorted(rng.sample(range(len(items)), cap))
            return OrderedDict(items[i] for i in idx)
        return OrderedDict(items[:cap])  # "first": registration order


def encrypted_policy() -> SelectionPolicy:
    """Exactly-two-worker demo policy (fc.py:397-407)."""
    return SelectionPolicy(lower_bound=LOWER_BOUND_ENC, upper_bound=UPPER_BOUND_ENC, policy="first")
