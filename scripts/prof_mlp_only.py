import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from colearn_federated_learning_b200 import ops
from colearn_federated_learning_b200.models import MLP, flatten_params
dev = torch.device("cuda", 0)
m = MLP(); spec = m.spec
theta = flatten_params(m).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.rand(n, 10, device=dev); y = (torch.rand(n, 1, device=dev) > 0.5).float()
perm = ops.device_permutation(n, 1, 0, dev)
out = torch.empty_like(theta)
descs = ops.build_client_descs([ops.ClientTask(x=x, y=y, theta_in=theta, theta_out=out, perm=perm)], dev)
for _ in range(3):
    ops.mlp_local_sgd_multi(spec.dims, "none", descs, 1, 1, 0.01, 1, -1, "xent")
torch.cuda.synchronize()
