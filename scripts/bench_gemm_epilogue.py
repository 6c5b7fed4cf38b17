#!/usr/bin/env python
"""Row-per-thread vs line-coalesced ("staged") GEMM epilogue, per epilogue mode, at the wide-MLP layer shape
(batch 1024, 4096 x 4096 layers = BASELINE config 5) and at 4096^3.  CUDA events, L2 flushed between iterations.

    python scripts/bench_gemm_epilogue.py > gpurun_out/r2_gemm_epilogue.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from colearn_federated_learning_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def case(m, n, k):
    bf = torch.bfloat16
    a, b = (torch.randn(m, k, device=dev) * 0.1).to(bf), (torch.randn(n, k, device=dev) * 0.1).to(bf)
    bias = torch.randn(n, device=dev)
    out, out_t = torch.empty(m, n, device=dev, dtype=bf), torch.empty(n, m, device=dev, dtype=bf)
    mask = torch.randn(m, n, device=dev).to(bf)
    cs = torch.empty(m // 32, n, device=dev)
    master, shadow = torch.randn(m, n, device=dev), torch.empty(m, n, device=dev, dtype=bf)
    modes = {
        "forward (bias+relu, bf16 + transposed)": dict(bias=bias, relu=True, out_bf16=out, out_bf16_t=out_t),
        "dgrad (relu mask, bf16 + transposed, colsum)": dict(relu_mask=mask, out_bf16=out, out_bf16_t=out_t, colsum=cs),
        "wgrad (fused SGD: fp32 master + bf16 shadow)": dict(sgd_master=master, sgd_lr=1e-4, sgd_shadow=shadow),
        "plain bf16 output": dict(out_bf16=out),
    }
    rows = []
    for name, kw in modes.items():
        row = {"shape": [m, n, k], "mode": name, "gflop": 2e-9 * m * n * k}
        for staged in (False, True):
            ms = timeit(lambda: ops.gemm_bf16(a, b, staged=staged, **kw))
            row["staged_ms" if staged else "row_per_thread_ms"] = round(ms, 4)
        row["speedup"] = round(row["row_per_thread_ms"] / row["staged_ms"], 3)
        rows.append(row)
    return rows


if __name__ == "__main__":
    res = []
    for shape in ((1024, 4096, 4096), (4096, 4096, 1024), (4096, 4096, 4096)):   # fwd/dgrad shape, wgrad shape, square
        res += case(*shape)
    print(json.dumps({"gemm_epilogue": res}))
