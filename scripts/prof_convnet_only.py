"""A few eager ResNet-18 steps on the repo's conv path, for `ncu -k regex:...` captures and launch-time lists.

    python scripts/prof_convnet_only.py [steps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colearn_federated_learning_b200.fl.convnet import ConvNetTrainer  # noqa: E402
from colearn_federated_learning_b200.models.registry import flatten_params  # noqa: E402
from colearn_federated_learning_b200.models.resnet import ResNet18  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ResNet18(10).to(dev)
flat = flatten_params(net)
tr = ConvNetTrainer(net, dev, 128, (32, 32))
tr.load(flat, net)
x = torch.randn(128, 3, 32, 32, device=dev)
y = torch.randint(0, 10, (128,), device=dev)
for _ in range(steps):
    loss = tr.step(x, y, 0.01)
tr.store(flat, net)
torch.cuda.synchronize()
print("loss", float(loss), "launches", tr.launches)
