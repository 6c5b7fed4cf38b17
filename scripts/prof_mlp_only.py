"""A few launches of the persistent local-SGD kernel (default variant of the net) for `ncu -k regex:mlp_local_sgd_kernel_v2`.

    python scripts/prof_mlp_only.py [samples] [mlp|ffnn]

The client is set up like the engine sets it up: the kernel makes its own keyed sample order (perm_seed + scratch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from colearn_federated_learning_b200 import ops
from colearn_federated_learning_b200.models import FFNN, MLP, flatten_params

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ffnn = len(sys.argv) > 2 and sys.argv[2] == "ffnn"
m = FFNN() if ffnn else MLP()
spec = m.spec
theta = flatten_params(m).to(dev)
x = torch.rand(n, 10, device=dev)
y = (torch.rand(n, 1, device=dev) > 0.5).float()
scratch = torch.empty(n, dtype=torch.int32, device=dev)
out = torch.empty_like(theta)
descs = ops.build_client_descs([ops.ClientTask(x=x, y=y, theta_in=theta, theta_out=out, perm_seed=12345, perm_row0=0, perm_scratch=scratch)], dev)
for _ in range(3):
    ops.mlp_local_sgd_multi(spec.dims, spec.out_activation, descs, 1, 1, 0.01, 1, -1, "sse" if ffnn else "xent")
torch.cuda.synchronize()
