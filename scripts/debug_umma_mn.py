#!/usr/bin/env python
"""If tests/test_zz_round2_gpu.py::test_gemm_mn_major_operands fails on the first GPU run: try the plausible
(LBO, SBO) readings of the MN-major shared-memory descriptor in one call (each in its own process, because the
binding reads COLEARN_UMMA_MN_LBO / COLEARN_UMMA_MN_SBO once) and print the max error of each.

    python scripts/debug_umma_mn.py            # parent: spawns one child per candidate
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CANDIDATES = [(0, 0), (8192, 1024), (1024, 8192), (8192, 128), (128, 1024), (1024, 1024), (16, 1024), (1024, 2048), (2048, 1024)]


def child() -> None:
    import torch

    from colearn_federated_learning_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    res = []
    for kdim, a_cols, m, n in [(64, 128, 128, 128), (64, 64, 128, 128), (256, 128, 128, 256)]:
        a = torch.randn(kdim, a_cols, device=dev).to(torch.bfloat16)
        b = torch.randn(kdim, n, device=dev).to(torch.bfloat16)
        out = torch.zeros(m, n, device=dev)
        ops.gemm_bf16(a, b, mn_m=m, out_f32=out)
        torch.cuda.synchronize()
        want = torch.zeros(m, n, device=dev)
        want[:a_cols] = a.float().t() @ b.float()
        res.append(float((out - want).abs().max()))
    print("max_abs_err", res, flush=True)


if __name__ == "__main__":
    if os.environ.get("_UMMA_CHILD") == "1":
        child()
        sys.exit(0)
    for lbo, sbo in CANDIDATES:
        env = dict(os.environ, _UMMA_CHILD="1", COLEARN_UMMA_MN_LBO=str(lbo), COLEARN_UMMA_MN_SBO=str(sbo))
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=120)
            tail = (p.stdout.strip().splitlines() or [p.stderr.strip()[-300:]])[-1]
        except subprocess.TimeoutExpired:
            tail = "TIMEOUT"
        print(f"LBO={lbo:5d} SBO={sbo:5d}: {tail}", flush=True)
