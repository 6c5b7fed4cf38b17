"""FL algorithms: FedAvg, local / remote / encrypted trainers, evaluation, inference."""
from .fedavg import federated_avg, federated_avg_flat, normalized_weights  # noqa: F401
from .trainer import FitConfig, local_fit, torch_fit, make_perm, resolve_loss  # noqa: F401
from .evaluate import evaluate, predict, forward_flat  # noqa: F401
