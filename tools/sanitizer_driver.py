#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck) over the kernels that changed in round 2: streaming scorer
(dynamic chunks + merge32 trees), mid-batch tree merge, exclusion filter, copy kernel, certified fp32 search with the
two-list cover + fused final form.  Sizes are tiny (the sanitizer runs kernels 10-100x slower).  Every result is compared
with the CPU oracle, so the script is also a quick stand-alone check where compute-sanitizer is not available (it is closed
on the round-2 GPU pool).

    compute-sanitizer --tool racecheck python tools/sanitizer_driver.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
orc = importlib.import_module("oracle.flat_oracle")
rng = np.random.default_rng(0)


def check(idx, xb, xq, k, metric, **kw):
    D, I = idx.search(xq, k, **kw)
    ref = orc.FlatIndexOracle(xb.shape[1], metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k)
    assert (I == Ir).all() and (D == Dr).all()


# streaming scorer: 1..4 queries, k = 15 (LIST1) and k = 40 (LIST4), host and device buffers
xb = rng.integers(-2, 3, size=(20_000, 128)).astype(np.float32)
for store in ("bf16", "f32"):
    idx = pkg.FlatIndex(128, pkg.METRIC_L2, store, device=0)
    idx.add(xb)
    for nq in (1, 3):
        xq = rng.integers(-2, 3, size=(nq, 128)).astype(np.float32)
        check(idx, xb, xq, 15, pkg.METRIC_L2)
        check(idx, xb, xq, 40, pkg.METRIC_L2)
        D, I = idx.search(torch.from_numpy(xq).cuda(), 15)
    idx.close()
# mid-batch tree merge (>= 8 lists) + tensor-core scorer with 10-entry lists
xb = rng.integers(-2, 3, size=(40_000, 64)).astype(np.float32)
xq = rng.integers(-2, 3, size=(200, 64)).astype(np.float32)
idx = pkg.FlatIndex(64, pkg.METRIC_IP, "bf16", device=0)
idx.add(xb)
idx.set_option("tc_chunks", 8)
check(idx, xb, xq, 10, pkg.METRIC_IP, algo="tc")
check(idx, xb, xq, 15, pkg.METRIC_IP, algo="tc")
# copy kernel
a = torch.arange(100_003, dtype=torch.uint8, device="cuda:0")
b = torch.zeros_like(a)
idx.copy_async(b[1:], a[1:])
c = torch.randn(4096, 33, device="cuda:0"); d = torch.empty_like(c)
idx.copy_async(d, c)
assert torch.equal(d, c) and torch.equal(b[1:], a[1:])
idx.close()
# certified fp32 search: tier 1 with the two-list cover (>= 262144 rows), re-rank with fused final form, tier-2 tail
xb = rng.standard_normal((262_144, 64)).astype(np.float32)
xq = rng.standard_normal((130, 64)).astype(np.float32)
idx = pkg.FlatIndex(64, pkg.METRIC_L2, "f32", device=0)
idx.add(xb)
D, I = idx.search(xq, 10)
assert idx.last_tier1[0] == 130
# exclusion filter + gather through the wrapper
class Cfg:
    vector_db_path = "/tmp/rdb_sanitizer"; vector_db_index_type = "L2"; top_k = 5; db_dtype = "f32"
vdb = pkg.VectorDatabase(Cfg())
vdb.index, vdb._cosine = idx, False
vdb.vector_paths = [f"p{i % 5000}" for i in range(262_144)]; vdb.vector_labels = [i & 1 for i in range(262_144)]
vec, lbl = pkg.retrieve_similar_vectors(vdb, torch.from_numpy(xq).cuda(), 5, query_paths=[f"/x/p{i}" for i in range(130)])
torch.cuda.synchronize()
print("sanitizer driver ok", tuple(vec.shape))
