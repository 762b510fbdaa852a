#!/usr/bin/env python
"""Measure the non-headline BASELINE.json configs (C1, C2, C4) and the ingest kernel on one B200.
Not a bench line -- writes JSON lines for profiles/.  Usage: python tools/bench_configs.py [c1 c2 c4 ingest ...]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")
orc = importlib.import_module("oracle.flat_oracle")
import torch  # noqa: E402

PEAK_HBM = 6547.8
try:
    PEAK_HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
dev = torch.device("cuda", 0)


def gen(n, d, seed, normalize=False):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn((n, d), generator=g, device=dev)
    return torch.nn.functional.normalize(x, dim=1) if normalize else x


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def parity(idx, xq_np, D, I, metric, store, cos, nsub, tol, k):
    """oracle on the first nsub queries against the STORED rows (read back), tolerance comparator."""
    n = idx.ntotal
    xb = idx.reconstruct_batch(torch.arange(n, device=dev)).cpu().numpy()
    ref = orc.FlatIndexOracle(idx.d, metric, store="f32")
    ref.add(xb)
    qn = orc.maybe_normalize(xq_np[:nsub], cos)
    if store != "f32":
        qn = orc.round_bf16(qn)
    Dr, Ir = ref.search(qn, k + 8, direct=False)
    scale = float((qn * qn).sum(1).max() + (xb * xb).sum(1).max())
    floor = (2e-6 if store == "f32" else 1e-4) * scale if metric == pkg.METRIC_L2 else 1e-6
    return orc.compare_topk(D[:nsub], I[:nsub], Dr, Ir, lambda ids: ref.exact_scores(qn, ids), metric, tol, floor)


def c1():
    N, Dm, Q, k = 20000, 768, 1000, 10
    xb, xq = gen(N, Dm, 1234).cpu().numpy(), gen(Q, Dm, 5678).cpu().numpy()
    for store in ("f32", "bf16"):
        idx = pkg.FlatIndex(Dm, pkg.METRIC_IP, store)
        idx.add(xb, normalize=True)
        dt, (D, I) = timed(lambda: idx.search(xq, k, normalize=True), 20)
        st = parity(idx, xq, D, I, pkg.METRIC_IP, store, True, Q, 1e-5 if store == "f32" else 1e-3, k)
        emit(config="C1 20k x 768, 1k queries, k=10 cosine, host numpy in/out", store=store, ms=dt * 1e3,
             qps=Q / dt, scorer=idx.last_kernel_ms()[1], kernel_ms=idx.last_kernel_ms()[0], parity=st)
    torch.set_num_threads(os.cpu_count())
    xbn = orc.maybe_normalize(xb, True)
    t0 = time.perf_counter()
    for _ in range(5):
        orc.torch_cpu_flat_search(xbn, orc.maybe_normalize(xq, True), k, orc.METRIC_IP)
    dt = (time.perf_counter() - t0) / 5
    emit(config="C1 CPU port (torch-CPU sgemm+topk)", cores=os.cpu_count(), ms=dt * 1e3, qps=Q / dt)


def c2():
    N, Dm, Q, k = 1_000_000, 768, 10000, 10
    xq = gen(Q, Dm, 5678)
    for metric, cos, name in ((pkg.METRIC_L2, False, "L2"), (pkg.METRIC_IP, True, "cosine")):
        idx = pkg.FlatIndex(Dm, metric, "f32")
        idx.reserve(N)
        for c in range(4):
            idx.add(gen(N // 4, Dm, 1234 + c), normalize=cos)
        dt, (D, I) = timed(lambda: idx.search(xq, k, normalize=cos), 3, warm=3)
        reps = []
        for _ in range(7):                                     # per-search wall times (each search ends synchronised)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            idx.search(xq, k, normalize=cos)
            torch.cuda.synchronize()
            reps.append((time.perf_counter() - t0) * 1e3)
        kms, scorer, ns = idx.last_kernel_ms()
        st = parity(idx, xq.cpu().numpy(), D.cpu().numpy(), I.cpu().numpy(), metric, "f32", cos, 128, 1e-5, k)
        emit(config=f"C2 1M x 768 fp32, 10k queries, k=10 {name}, device in/out", ms=dt * 1e3, qps=Q / dt,
             scorer=scorer, kernel_ms=kms, tflops=2.0 * Q * N * Dm / (kms * 1e-3) / 1e12, parity_128q=st,
             uncertified_queries=idx.last_uncertified, tier1=idx.last_tier1, per_search_ms=sorted(reps),
             qps_median=Q / (sorted(reps)[len(reps) // 2] * 1e-3))
        dt, _ = timed(lambda: idx.search(xq, k, normalize=cos, algo="simt"), 2, warm=1)
        emit(config=f"C2 (exact CUDA-core kernel only) {name}", ms=dt * 1e3, qps=Q / dt,
             tflops=2.0 * Q * N * Dm / dt / 1e12)
        idx.close()


def c4():
    N, Dm, k = 1_000_000, 768, 15
    for store in ("bf16", "f32"):
        idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, store)
        idx.reserve(N)
        for c in range(4):
            idx.add(gen(N // 4, Dm, 1234 + c))
        xq = gen(2000, Dm, 5678).cpu().numpy()
        for nq in (1, 2):
            lat, kms = [], []
            for i in range(0, 600, nq):
                t0 = time.perf_counter()
                idx.search(xq[i:i + nq], k)                      # host in, host out: H2D + kernels + D2H + sync
                lat.append((time.perf_counter() - t0) * 1e3)
                kms.append(idx.last_kernel_ms()[0])
            lat, kms = np.sort(lat[20:]), np.sort(kms[20:])
            bytes_db = N * Dm * (2 if store == "bf16" else 4)
            emit(config=f"C4 1M x 768 {store}, batch-{nq} streaming, k=15 L2, host in/out",
                 p50_ms=float(lat[len(lat) // 2]), p99_ms=float(lat[int(len(lat) * 0.99)]),
                 kernel_p50_ms=float(kms[len(kms) // 2]), scorer=idx.last_kernel_ms()[1],
                 hbm_floor_ms=bytes_db / (PEAK_HBM * 1e9) * 1e3,
                 kernel_hbm_frac=bytes_db / (float(kms[len(kms) // 2]) * 1e-3) / 1e9 / PEAK_HBM)
        idx.close()


def ingest():
    N, Dm = 2_000_000, 768
    x = gen(N, Dm, 1)
    for store, cos in (("bf16", True), ("bf16", False), ("f32", True)):
        idx = pkg.FlatIndex(Dm, pkg.METRIC_IP, store)
        idx.reserve(N)
        idx.add(x[:1000], normalize=cos)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx.add(x[1000:], normalize=cos)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows = N - 1000
        rd = rows * Dm * 4 * (2 if cos else 1)                   # normalise reads the row twice (2nd pass hits L2/L1)
        wr = rows * (Dm * 2 + 4) if store == "bf16" else rows * (Dm * 4 + Dm * 2 * 2 + 4)
        alg = rows * Dm * 4 + wr                                 # algorithmic: read once + writes
        emit(config=f"ingest 2M x 768 fp32 -> {store}, normalize={cos}, device tensor in", ms=ms,
             rows_per_s=rows / (ms * 1e-3), algorithmic_GBs=alg / (ms * 1e-3) / 1e9,
             frac_of_hbm_peak=alg / (ms * 1e-3) / 1e9 / PEAK_HBM)
        idx.close()


def torch_recall(idx, q_dev, I_ours, k, normalize, nsub=128, metric_ip=True):
    """recall@k vs a torch fp32 brute force over the stored rows (checker for sizes the CPU oracle cannot hold)."""
    n = idx.ntotal
    qs = q_dev[:nsub]
    if normalize:
        qs = torch.nn.functional.normalize(qs, dim=1, eps=1e-12)
    qs = qs.to(torch.bfloat16).to(torch.float32)
    bv = torch.full((nsub, k), float("-inf"), device=dev)
    bi = torch.full((nsub, k), -1, dtype=torch.int64, device=dev)
    step = 1_000_000
    for s0 in range(0, n, step):
        e = min(n, s0 + step)
        rows = idx.reconstruct_batch(torch.arange(s0, e, device=dev))
        sc = qs @ rows.T
        if not metric_ip:
            sc = 2 * sc - (rows * rows).sum(1)[None, :]
        v, i = torch.topk(sc, k, dim=1)
        cv, ci = torch.cat([bv, v], 1), torch.cat([bi, i + s0], 1)
        bv, sel = torch.topk(cv, k, dim=1)
        bi = torch.gather(ci, 1, sel)
    ours = I_ours[:nsub]
    return (ours.unsqueeze(2) == bi.unsqueeze(1)).any(2).float().mean().item()


def c5():
    """One GPU's share of C5: 100M x 256 bf16 over 8 GPUs = 12.5M rows per GPU, k = 100."""
    N, Dm, k = 12_500_000, 256, 100
    idx = pkg.FlatIndex(Dm, pkg.METRIC_IP, "bf16")
    idx.reserve(N)
    for c in range(25):
        idx.add(gen(N // 25, Dm, 1234 + c), normalize=True)
    for Q in (16384, 1):
        xq = gen(Q, Dm, 5678)
        dt, (D, I) = timed(lambda: idx.search(xq, k, normalize=True), 3 if Q > 1 else 50, warm=2)
        kms, scorer, ns = idx.last_kernel_ms()
        rec = torch_recall(idx, xq, I, k, True, nsub=min(Q, 64))
        emit(config=f"C5 per-GPU share: 12.5M x 256 bf16, {Q} queries, k=100 cosine, device in/out", ms=dt * 1e3,
             qps_per_gpu=Q / dt, scorer=scorer, splits=ns, kernel_ms=kms,
             tflops=2.0 * Q * N * Dm / (kms * 1e-3) / 1e12, frac_of_sustained_bf16=2.0 * Q * N * Dm / (kms * 1e-3) / 1e12 / 1338.4,
             hbm_frac=N * Dm * 2 / (kms * 1e-3) / 1e9 / PEAK_HBM, recall_vs_torch_fp32=rec)
    # k = 10 on the same shard for comparison (register-resident list)
    xq = gen(16384, Dm, 5678)
    dt, (D, I) = timed(lambda: idx.search(xq, 10, normalize=True), 3, warm=2)
    kms, scorer, ns = idx.last_kernel_ms()
    emit(config="C5 shard, k=10 for comparison", ms=dt * 1e3, kernel_ms=kms, splits=ns,
         tflops=2.0 * 16384 * N * Dm / (kms * 1e-3) / 1e12)
    idx.close()


def refscale():
    """Reference scale R: N = 25 423, D = 5376 (7 x 768 TPP), one training batch of Q = 256 queries,
    K = 5 neighbours of the K + 10 searched, L2 fp32 -- the whole caller step pipeline.py:449-532 (search + rank-ordered
    self-exclusion + gather of [B, K, D] neighbours + labels), device tensors in and out, against the caller restated
    on the CPU port (oracle index + the reference's Python double loop)."""
    import tempfile

    class Cfg:
        vector_db_path = tempfile.mkdtemp()
        vector_db_index_type = "L2"
        top_k = 5
        use_float16 = False
        vector_add_batch_size = 10000

    N, Dm, Q, K = 25423, 5376, 256, 5
    xb = gen(N, Dm, 1234).cpu().numpy()
    paths = [f"/data/spk{i % 97}/utt_{i:06d}.wav" for i in range(N)]
    labels = [int(i % 2) for i in range(N)]
    q = torch.from_numpy(xb[:Q] + 0.05 * gen(Q, Dm, 5678).cpu().numpy()).cuda()
    qpaths = paths[:Q]                                         # the batch's own files are in the DB: exclusion matters
    for store in ("f32", "bf16"):
        cfg = Cfg()
        cfg.db_dtype = store
        vdb = pkg.VectorDatabase(cfg)
        vdb.add_vectors(xb, paths, labels, {"speaker_id": [p.split("/")[2] for p in paths]})

        def step():
            out = pkg.retrieve_similar_vectors(vdb, q, K, query_paths=qpaths, exclude_self=True, return_distances=True)
            return out
        dt, out = timed(step, 20, warm=3)
        emit(config=f"R: retrieve_similar_vectors, N=25423 D=5376 Q=256 K=5(+10) L2, store {store}, device in/out",
             ms=dt * 1e3, batches_per_s=1.0 / dt, scorer=vdb.index.last_kernel_ms()[1],
             search_kernel_ms=vdb.index.last_kernel_ms()[0])
        if store == "f32":
            ours = [t.cpu().numpy() for t in out]
            # the UNMODIFIED reference caller (per-neighbour index.reconstruct loop, restated in the oracle) on top of
            # this index: module-replacement integration
            qn_ = q.cpu().numpy()
            for _ in range(2):
                orc.retrieve_similar_vectors_oracle(vdb, qn_, K, Dm, query_paths=qpaths, exclude_self=True)
            t0 = time.perf_counter()
            for _ in range(5):
                v_, l_, p__, d_ = orc.retrieve_similar_vectors_oracle(vdb, qn_, K, Dm, query_paths=qpaths, exclude_self=True)
            dtu = (time.perf_counter() - t0) / 5
            emit(config="R: unmodified reference caller loop (search_batch + 256 x <=15 index.reconstruct calls) on this index, f32",
                 ms=dtu * 1e3, batches_per_s=1.0 / dtu,
                 identical_to_device_path=bool(np.array_equal(ours[0], v_) and np.array_equal(ours[1], l_)))
        vdb.index.close()
    torch.set_num_threads(os.cpu_count())
    ovdb = orc.OracleVectorDatabase(Cfg())
    ovdb.add_vectors(xb, paths, labels, {"speaker_id": [p.split("/")[2] for p in paths]})
    qn = q.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(3):
        v, l, p_, d = orc.retrieve_similar_vectors_oracle(ovdb, qn, K, Dm, query_paths=qpaths, exclude_self=True)
    dt = (time.perf_counter() - t0) / 3
    same = bool(np.array_equal(ours[1], l)) and bool(np.allclose(ours[0], v, rtol=0, atol=0))
    emit(config="R: the same caller step on the CPU port (oracle index, reference's Python loop)", cores=os.cpu_count(),
         ms=dt * 1e3, batches_per_s=1.0 / dt, neighbours_and_labels_identical_to_gpu_f32=same)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "ingest", "c4", "c2"]
    for w in which:
        globals()[w]()
