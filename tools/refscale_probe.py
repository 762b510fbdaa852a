#!/usr/bin/env python
"""The reference's own shapes (25 423 stored segments x 5376 features, batches of 256, top_k = 5 + 10; pipeline.py:449-532)
through retrieve_similar_vectors, device in/out: ms per batch for the fp32 (exact rows) and bf16 stores.  JSON lines."""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
dev = torch.device("cuda", 0)
N, D, B = 25_423, 5376, 256
g = torch.Generator(device=dev)
g.manual_seed(99)
MODE = sys.argv[1] if len(sys.argv) > 1 else "gauss"
if MODE == "clustered":
    # segments of the same utterance / speaker sit close together: 1000 centres, rows = centre + 0.35 * noise, queries =
    # fresh points of 256 of the clusters (the shape of pooled speech embeddings; iid Gaussian rows are the worst case
    # for the certificate because the neighbours are barely closer than everything else)
    cen = torch.randn((1000, D), generator=g, device=dev)
    xb = cen[torch.randint(0, 1000, (N,), generator=g, device=dev)] + 0.35 * torch.randn((N, D), generator=g, device=dev)
    xq = cen[torch.randint(0, 1000, (B,), generator=g, device=dev)] + 0.35 * torch.randn((B, D), generator=g, device=dev)
else:
    xb = torch.randn((N, D), generator=g, device=dev)
    xq = torch.randn((B, D), generator=g, device=dev)
qp = [f"/data/train/p{i}" for i in range(B)]
for dtype in ("f32", "bf16"):
    for tier1 in ((1, 0) if dtype == "f32" else (1,)):
        class Cfg:
            vector_db_path = "/tmp/rdb_refscale_probe"; vector_db_index_type = "L2"; top_k = 5; db_dtype = dtype
        vdb = pkg.VectorDatabase(Cfg())
        vdb.create_index(D)
        vdb.add_vectors(xb, [f"p{i}" for i in range(N)], [i & 1 for i in range(N)], {})
        vdb.index.set_option("tier1", tier1)
        for _ in range(5):
            pkg.retrieve_similar_vectors(vdb, xq, 5, query_paths=qp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            vec, lbl = pkg.retrieve_similar_vectors(vdb, xq, 5, query_paths=qp)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / 50
        print(json.dumps({"data": MODE, "config": f"retrieve_similar_vectors {N}x{D} {dtype}, {B} queries, K=5(+10), L2, device in/out",
                          "tier1": tier1, "ms_per_batch": ms, "search_kernel_ms": vdb.index.last_kernel_ms()[0],
                          "last_tier1": list(vdb.index.last_tier1) if dtype == "f32" else None}), flush=True)
        vdb.cleanup_gpu_resources()
