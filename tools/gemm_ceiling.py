#!/usr/bin/env python
"""Library-GEMM yardstick for the headline kernel: cuBLAS bf16 Q x Y^T at the C3 contraction shape (K = 768), output
written to HBM (which the fused kernel never does), back to back for ~4 s so the power cap settles.  Prints one JSON
line.  Profiling aid only -- not a product path."""
import json
import sys
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    nq, n = 16384, 131072
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn((nq, D), generator=g, device=dev), dim=1).to(torch.bfloat16)
    y = torch.nn.functional.normalize(torch.randn((n, D), generator=g, device=dev), dim=1).to(torch.bfloat16)
    out = torch.empty((nq, n), dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        torch.matmul(q, y.T, out=out)
    torch.cuda.synchronize()
    flops = 2.0 * nq * n * D
    # burst: best of 10 single launches
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(q, y.T, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~4 s
    reps = max(10, int(4000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(reps):
        torch.matmul(q, y.T, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"what": f"cuBLAS bf16 [{nq},{D}] x [{n},{D}]^T -> bf16 in HBM", "burst_tflops": flops / best / 1e9,
                      "sustained_tflops": flops / ms / 1e9, "reps": reps, "wall_s": time.time() - t0}))


if __name__ == "__main__":
    main()
