#!/usr/bin/env python
"""Time the k > 128 path (dense keys + radix select) on one B200; JSON lines for profiles/.  Not a bench line.
Usage: python tools/bench_large_k.py"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
import torch  # noqa: E402

dev = torch.device("cuda", 0)


def gen(n, d, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((n, d), generator=g, device=dev), dim=1)


def run(N, D, Q, k, store, reps=5):
    idx = pkg.FlatIndex(D, pkg.METRIC_IP, store)
    for s in range(0, N, 1 << 18):
        idx.add(gen(min(1 << 18, N - s), D, 1234 + s))
    xq = gen(Q, D, 5678)
    for _ in range(2):
        idx.search(xq, k)
    torch.cuda.synchronize()
    l0 = idx.launch_count
    per = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Dd, Ii = idx.search(xq, k)
        torch.cuda.synchronize()
        per.append(time.perf_counter() - t0)
    dt = sorted(per)[len(per) // 2]
    # sanity: brute force on a few queries with torch (fp32 matmul on the stored rows)
    xb = idx.reconstruct_batch(torch.arange(N, device=dev))
    qs = xq[:8].to(torch.bfloat16).float() if store != "f32" else xq[:8]
    ref = torch.topk(qs @ xb.T, k, dim=1).indices
    rec = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ref.cpu(), Ii[:8].cpu())) / (8 * k)
    print(json.dumps({"what": "large_k", "N": N, "D": D, "Q": Q, "k": k, "store": store, "ms": dt * 1e3,
                      "per_search_ms": [round(p * 1e3, 3) for p in per], "scorer": idx.last_kernel_ms()[1], "qps": Q / dt, "tflops": 2.0 * N * D * Q / dt / 1e12, "recall_vs_torch_8q": rec,
                      "launches_per_search": (idx.launch_count - l0) / reps}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
        sys.exit(0)
    run(1_000_000, 768, 1000, 1000, "bf16")
    run(1_000_000, 768, 1000, 200, "f32")
    run(25_423, 5376, 256, 500, "f32")
    run(4_000_000, 256, 512, 2048, "bf16")
