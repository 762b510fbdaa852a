#!/usr/bin/env python
"""Where the batch-1 host-to-host latency goes (C4: 1M x 768, k = 15): the scorer kernel (CUDA events), the C-ABI call
with host buffers (rdb_search timed around the ctypes call), FlatIndex.search, VectorDatabase.search.  JSON lines.
Profiling aid only.   python tools/latency_probe.py [bf16|f32] [N]"""
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")

store = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
D, K, REPS = 768, 15, 5000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
idx = pkg.FlatIndex(D, pkg.METRIC_IP, store)
idx.reserve(N)
for c in range(0, N, 250_000):
    idx.add(torch.randn((min(250_000, N - c), D), generator=g, device=dev), normalize=True)
qs = np.random.default_rng(2).standard_normal((REPS, D)).astype(np.float32)


def pct(v):
    v = sorted(v)
    return {"p50_ms": v[len(v) // 2] * 1e3, "p99_ms": v[int(len(v) * 0.99)] * 1e3, "min_ms": v[0] * 1e3}


class Cfg:
    vector_db_path = "/tmp/rdb_latency_probe"
    vector_db_index_type = "IP"
    top_k = 5
    db_dtype = store


vdb = pkg.VectorDatabase(Cfg())
vdb.index, vdb._cosine = idx, True
lib = idx._lib
Dd = np.empty((1, K), np.float32)
Ii = np.empty((1, K), np.int64)
vp = ctypes.c_void_p
for name, fn in (
        ("c_abi rdb_search(host)", lambda q: lib.rdb_search(idx._h, q.ctypes.data_as(vp), 1, K, 0, 1, Dd.ctypes.data_as(vp),
                                                            Ii.ctypes.data_as(vp), None)),
        ("FlatIndex.search", lambda q: idx.search(q.reshape(1, -1), K, normalize=True)),
        ("VectorDatabase.search", lambda q: vdb.search(q, k=K))):
    for i in range(300):
        fn(qs[i])
    lat, kms = [], []
    for i in range(REPS):
        t0 = time.perf_counter()
        fn(qs[i])
        lat.append(time.perf_counter() - t0)
        if i % 50 == 0:
            kms.append(idx.last_kernel_ms()[0])
    out = {"what": name, "store": store, "N": N, "kernel_ms_median": float(np.median(kms))}
    out.update(pct(lat))
    floor = N * D * (2 if store != "f32" else 4) / 6547.8e9 * 1e3
    out["hbm_floor_ms"] = floor
    out["p50_frac_of_floor"] = floor / out["p50_ms"]
    print(json.dumps(out), flush=True)
