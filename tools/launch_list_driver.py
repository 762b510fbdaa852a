#!/usr/bin/env python
"""Drive three searches of one BASELINE config (c1 | c2 | c5) so that
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` lists every kernel of a search step: how the step
splits between the scorer and the auxiliary kernels (query prep, merge, re-rank, pivot).  Profiling aid only."""
import importlib, sys, os, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
which = sys.argv[1]
if which == "c2":
    N, D, Q, k, store, met = 1000000, 768, 10000, 10, "f32", pkg.METRIC_L2
elif which == "c2cos":
    N, D, Q, k, store, met = 1000000, 768, 10000, 10, "f32", pkg.METRIC_IP
elif which == "largek":
    N, D, Q, k, store, met = 1000000, 768, 1000, 1000, "bf16", pkg.METRIC_IP
elif which.startswith("mid"):          # mid-batch regime (the reference's own batch size is 256): mid128 / mid256 / mid512
    N, D, Q, k, store, met = 1000000, 768, int(which[3:]), 15, "bf16", pkg.METRIC_IP
elif which == "refscale":             # the reference's own shapes: retrieve_similar_vectors (search + filter + gather), fp32 store
    N, D, Q, k, store, met = 25423, 5376, 256, 15, "f32", pkg.METRIC_L2
elif which == "c1":
    N, D, Q, k, store, met = 20000, 768, 1000, 10, "f32", pkg.METRIC_IP
else:
    N, D, Q, k, store, met = 12500000, 256, 16384, 100, "bf16", pkg.METRIC_IP
idx = pkg.FlatIndex(D, met, store); idx.reserve(N)
for c in range(0, N, 500000):
    idx.add(torch.randn((min(500000, N - c), D), generator=g, device=dev), normalize=(met == pkg.METRIC_IP))
q = torch.randn((Q, D), generator=g, device=dev)
if which == "refscale":
    class _Cfg:
        vector_db_path = "/tmp/rdb_launch_list"; vector_db_index_type = "L2"; top_k = 5; db_dtype = store
    vdb = pkg.VectorDatabase(_Cfg())
    vdb.index, vdb._cosine = idx, False
    vdb.vector_paths = [f"p{i}" for i in range(N)]; vdb.vector_labels = [i & 1 for i in range(N)]
    qp = [f"/x/p{i}" for i in range(Q)]
    for _ in range(3):
        pkg.retrieve_similar_vectors(vdb, q, 5, query_paths=qp)
    torch.cuda.synchronize()
    print(which, idx.last_kernel_ms())
    sys.exit(0)
for _ in range(3):
    idx.search(q, k, normalize=(met == pkg.METRIC_IP))
torch.cuda.synchronize()
print(which, idx.last_kernel_ms())
