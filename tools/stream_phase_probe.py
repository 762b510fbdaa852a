#!/usr/bin/env python
"""Where the time of ONE streaming-scorer launch goes (kernel 3, the batch-1 latency path): per-block globaltimer
stamps written by a RDB_PROFILING build (`make RDB_PROFILING=1 OBJDIR=build_prof OUT=../libradad_flat_prof.so`,
option "stream_prof").  Phases: start -> queries prepared -> rows streamed -> block lists merged -> (last block)
final merge -> results published.  JSON lines.   python tools/stream_phase_probe.py [bf16|f32] [N]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "radad-retrievalaugmenteddeepfakeaudiodetection_b200"
cabi = importlib.import_module(PKG + "._cabi")
cabi.LIB_PATH = os.path.join(ROOT, PKG, "libradad_flat_prof.so")
pkg = importlib.import_module(PKG)

store = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
D, K = 768, 15
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
idx = pkg.FlatIndex(D, pkg.METRIC_IP, store)
idx.reserve(N)
for c in range(0, N, 250_000):
    idx.add(torch.randn((min(250_000, N - c), D), generator=g, device=dev), normalize=True)
qs = np.random.default_rng(2).standard_normal((400, D)).astype(np.float32)
prof = torch.zeros((148, 8), dtype=torch.int64, device=dev)
idx.set_option("stream_prof", prof.data_ptr())
rows = []
for i in range(400):
    prof.zero_()
    torch.cuda.synchronize()
    idx.search(qs[i:i + 1], K, normalize=True)
    torch.cuda.synchronize()
    if i < 100:
        continue
    t = prof.cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    last = int(np.argmax(t[:, 5]))
    rows.append({
        "kernel_ms_events": idx.last_kernel_ms()[0],
        "start_spread_us": (t[:, 0].max() - t0) / 1e3,
        "prep_us_median": float(np.median(t[:, 1] - t[:, 0])) / 1e3,
        "stream_us_median": float(np.median(t[:, 2] - t[:, 1])) / 1e3,
        "stream_end_first_us": (t[:, 2].min() - t0) / 1e3,
        "stream_end_last_us": (t[:, 2].max() - t0) / 1e3,
        "block_merge_us_median": float(np.median(t[:, 3] - t[:, 2])) / 1e3,
        "last_block_wait_us": (t[last, 4] - t[last, 3]) / 1e3,
        "final_merge_us": (t[last, 5] - t[last, 4]) / 1e3,
        "total_us": (t[last, 5] - t0) / 1e3,
    })
out = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
out.update({"what": "stream kernel phases (median of 300 launches, globaltimer)", "store": store, "N": N, "D": D, "k": K})
print(json.dumps(out), flush=True)
