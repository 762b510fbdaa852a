#!/usr/bin/env python
"""C3 through the single-process drop-in: ONE VectorDatabase (config.db_devices = all GPUs of the box), host numpy
queries in, host numpy results out -- what an unmodified pipeline.py would see.  Prints one JSON line.
Not the driver's bench line (that is bench.py under torchrun); a measurement for profiles/."""
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

N = int(os.environ.get("RDB_BENCH_N", 10_000_000))
D = int(os.environ.get("RDB_BENCH_D", 768))
Q = int(os.environ.get("RDB_BENCH_Q", 65536))
K = int(os.environ.get("RDB_BENCH_K", 10))
STEPS = int(os.environ.get("RDB_BENCH_STEPS", 5))
DTYPE = os.environ.get("RDB_BENCH_DTYPE", "bf16")


class Cfg:
    vector_db_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"rdb_sp_{os.getpid()}")
    vector_db_index_type = "IP"
    top_k = K
    db_dtype = DTYPE
    db_devices = "all"


def main():
    G = torch.cuda.device_count()
    vdb = pkg.VectorDatabase(Cfg())
    vdb.create_index(D)
    idx = vdb.index
    idx.reserve(N)
    gen = torch.Generator(device="cuda:0")
    for c in range(0, N, 250_000):
        gen.manual_seed(1234 + c // 250_000)
        x = torch.randn((min(250_000, N - c), D), generator=gen, device="cuda:0")
        idx.add(x, normalize=True)
    gen.manual_seed(5678)
    q = torch.randn((Q, D), generator=gen, device="cuda:0").cpu().numpy()
    for _ in range(3):
        Dn, In = vdb.search_batch(q, k=K)
    for g in range(G):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(STEPS):
        Dn, In = vdb.search_batch(q, k=K)
    dt = (time.perf_counter() - t0) / STEPS
    # device-resident variant (queries already on GPU 0)
    qd = torch.from_numpy(q).cuda(0)
    for _ in range(2):
        idx.search(qd, K, normalize=True)
    torch.cuda.synchronize(0)
    t0 = time.perf_counter()
    for _ in range(STEPS):
        Dd, Id = idx.search(qd, K, normalize=True)
    torch.cuda.synchronize(0)
    dt_dev = (time.perf_counter() - t0) / STEPS
    phases = None
    if hasattr(idx, "phase_probe"):
        idx.phase_probe = True
        t0 = time.perf_counter()
        vdb.search_batch(q, k=K)
        probe_ms = (time.perf_counter() - t0) * 1e3
        phases = {"total_ms_with_probe": probe_ms, "per_gpu": idx.last_phases}
        idx.phase_probe = "events"
        t0 = time.perf_counter()
        vdb.search_batch(q, k=K)
        phases["events_probe"] = {"total_ms": (time.perf_counter() - t0) * 1e3, "per_gpu": idx.last_phases}
        idx.phase_probe = False
    print(json.dumps({"phases": phases}))
    print(json.dumps({"what": "single-process multi-GPU VectorDatabase.search_batch (host numpy in/out)",
                      "workload": f"{N}x{D} {DTYPE}, {Q} queries, k={K}, cosine", "n_gpus": G,
                      "shard_sizes": idx.shard_sizes if hasattr(idx, "shard_sizes") else [idx.ntotal],
                      "e2e_qps": Q / dt, "e2e_ms": dt * 1e3, "device_qps": Q / dt_dev, "device_ms": dt_dev * 1e3,
                      "kernel_ms_max": idx.last_kernel_ms()[0],
                      "ids_equal_host_vs_device": bool(np.array_equal(In, Id.cpu().numpy()))}))


if __name__ == "__main__":
    main()
