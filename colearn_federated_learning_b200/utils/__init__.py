"""Utilities: checkpoint/resume, round metrics, monitors, timing."""
from .checkpoint import (save_model, save_state_dict, load_or_init, load_meta, checkpoint_compatible,  # noqa: F401
                         DEFAULT_PATH)
from .metrics import RoundLogger  # noqa: F401
