import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from colearn_federated_learning_b200 import ops
dev = torch.device("cuda", 0)
m = n = k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a = (torch.randn(m, k, device=dev) * 0.1).to(torch.bfloat16)
b = (torch.randn(n, k, device=dev) * 0.1).to(torch.bfloat16)
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
master = torch.randn(m, n, device=dev)
shadow = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm_bf16(a, b, out_bf16=out)                                   # cta_group::2 (auto)
    ops.gemm_bf16(a, b, out_bf16=out, tile_n=256, cluster=1)            # single-CTA 128x256
    ops.gemm_bf16(a, b, sgd_master=master, sgd_lr=1e-3, sgd_shadow=shadow)   # wgrad epilogue (fused SGD)
torch.cuda.synchronize()
theta, slots = torch.zeros(1 << 24, device=dev), torch.randn(8, 1 << 24, device=dev)
w = torch.full((8,), 0.125, device=dev)
for _ in range(3):
    ops.fedavg_apply(theta, slots, w, 1.0)
    ops.sgd_step(theta, slots[0], 0.01)
torch.cuda.synchronize()
