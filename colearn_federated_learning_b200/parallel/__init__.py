"""Multi-GPU runtime: symmetric memory, fused NVLink collectives, the federated round engine."""
from .engine import FederatedEngine, RoundReport  # noqa: F401
from .launcher import init_distributed, shutdown, env_world  # noqa: F401
