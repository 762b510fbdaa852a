#!/usr/bin/env python
"""A/B of the pipelined host-buffer search (option "host_pipeline"): pageable numpy queries in, numpy results out,
through FlatIndex.search.  Writes JSON lines for profiles/.  Usage: python tools/host_pipeline_ab.py [c2 c3]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
import torch  # noqa: E402

dev = torch.device("cuda", 0)


def run(name, N, D, Q, k, store, reps):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    idx = pkg.FlatIndex(D, pkg.METRIC_IP, store)
    for s in range(0, N, 1 << 20):
        n = min(1 << 20, N - s)
        idx.add(torch.randn((n, D), generator=g, device=dev), normalize=True)
    xq = torch.randn((Q, D), generator=g, device=dev).cpu().numpy()     # pageable
    res = {}
    for rnd in range(2):
        for mode in (0, 1):
            idx.set_option("host_pipeline", mode)
            idx.search(xq, k, normalize=True)
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                Dd, I = idx.search(xq, k, normalize=True)
                ts.append((time.perf_counter() - t0) * 1e3)
            res.setdefault(mode, []).extend(ts)
            res[("I", mode)] = I
    same = bool((res[("I", 0)] == res[("I", 1)]).all())
    xq_d = torch.from_numpy(xq).to(dev)
    idx.search(xq_d, k, normalize=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        idx.search(xq_d, k, normalize=True)
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) * 1e3 / reps
    print(json.dumps({"config": name, "N": N, "D": D, "Q": Q, "k": k, "store": store,
                      "host_ms_unpipelined_median": float(np.median(res[0])), "host_ms_pipelined_median": float(np.median(res[1])),
                      "device_in_out_ms": dev_ms, "ids_equal": same, "query_bytes": int(xq.nbytes)}), flush=True)
    idx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3"]
    if "c1" in which:
        run("C1-like 20k x 768 bf16, 8192 host queries", 20000, 768, 8192, 10, "bf16", 10)
    if "c2" in which:
        run("C2 1M x 768 fp32, 10k host queries, k=10 cosine", 1_000_000, 768, 10000, 10, "f32", 8)
        run("C2-shape bf16 store", 1_000_000, 768, 10000, 10, "bf16", 8)
    if "c3" in which:
        run("C3 10M x 768 bf16, 65536 host queries, k=10 cosine", 10_000_000, 768, 65536, 10, "bf16", 3)
