#!/usr/bin/env python
"""Drive the HBM-bound kernels once each for an ncu capture: ingest (2M x 768 fp32 -> bf16, normalise) and the
batch-1 streaming search (C4: 1M x 768 bf16, k = 15; C5 shard: 12.5M x 256 bf16, k = 100), or -- `largek` -- the
k > 128 path (1M x 768 bf16, 256 queries, k = 1000: SelectDump scorer + select_dense_kernel).  Profiling aid only.

    ncu --set full --clock-control none --import-source on -k regex:select_dense -c 1 -o gpurun_out/select_dense \
        python tools/ncu_small_kernels.py largek"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    if which == "largek":
        N, D, k = 1_000_000, 768, 1000
    elif which == "c4":
        N, D, k = 1_000_000, 768, 15
    else:
        N, D, k = 12_500_000, 256, 100
    idx = pkg.FlatIndex(D, pkg.METRIC_L2 if which == "c4" else pkg.METRIC_IP, "bf16")
    idx.reserve(N)
    for c in range(0, N, 500_000):
        x = torch.randn((min(500_000, N - c), D), generator=g, device=dev)
        idx.add(x, normalize=(which != "c4"))                      # ingest kernel launches
    q = torch.randn((256 if which == "largek" else 1, D), generator=g, device=dev)
    for _ in range(5):
        idx.search(q, k, normalize=(which != "c4"))               # streaming search launches
    torch.cuda.synchronize()
    print(which, idx.last_kernel_ms())


if __name__ == "__main__":
    main()
