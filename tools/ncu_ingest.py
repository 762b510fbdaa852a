#!/usr/bin/env python
"""Drive the ingest kernel for an ncu capture: 1M x 768 fp32 device rows -> bf16 store, with (default) or without
(`plain`) the fused numpy-order normalisation.  Profiling aid only.

    ncu --set full --clock-control none --import-source on -k regex:ingest_rows -s 2 -c 1 -o gpurun_out/ingest \
        python tools/ncu_ingest.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
norm = not (len(sys.argv) > 1 and sys.argv[1] == "plain")
x = torch.randn((1_000_000, 768), generator=g, device=dev)
idx = pkg.FlatIndex(768, pkg.METRIC_IP, sys.argv[2] if len(sys.argv) > 2 else "bf16")
idx.reserve(4_000_000)
for _ in range(4):
    idx.add(x, normalize=norm)
torch.cuda.synchronize()
print("ingest", norm, idx.ntotal)
