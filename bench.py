#!/usr/bin/env python
"""bench.py -- headline benchmark of the RADAD retrieval hot path on B200.

Metric (BASELINE.json): QPS @ k=10 on a 10M x 768 database, 64k-query batch, cosine (IP + normalise),
bf16 storage, row-sharded over N GPUs (one process per GPU; NCCL all-gather of the per-shard top-k +
on-device merge).  A "step" is one pass of the hot path over the whole 65 536-query batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Prints ONE JSON line on rank 0.  Keys follow the driver's contract:
  value        whole-job QPS with the queries already resident in HBM when the timed region starts
  e2e          the same metric through the reference-facing call with HOST buffers (pinned), host<->device
               copies inside the timed region (N=1: VectorDatabase.search_batch(numpy); N>1: sharded search)
  roofline     dominant kernel (tcgen05 score+select): algorithmic FLOPs 2*Q*N_shard*D per launch / its average
               CUDA-event duration, against the measured sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline the oracle's torch-CPU restatement of the reference's faiss sgemm path on the host cores, on a
               bounded sample (rank 0, N=1 only); its neighbour ids are compared with the CUDA path's on the same rows
  secondary    the other BASELINE.json configs measured in the same run (N=1): C2 (1M x 768 fp32, 10k queries, L2 and
               cosine, exact-fp32 neighbours), C4 (batch-1 latency over 10 000 sequential host queries, bf16 and fp32),
               ingest GB/s with / without the fused normalisation -- each beside its roofline
  correctness  recall@10 of >= 4096 queries against a brute force over the stored rows (N>1: per-rank brute force over the
               local shard -> all-gather -> merge) and, at N>1, ids_equal_n1_subset: the N-GPU answer against the answer of
               ONE GPU holding the whole database
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "radad-retrievalaugmenteddeepfakeaudiodetection_b200"

N_DB = int(os.environ.get("RDB_BENCH_N", 10_000_000))
DIM = int(os.environ.get("RDB_BENCH_D", 768))
NQ = int(os.environ.get("RDB_BENCH_Q", 65536))
K = int(os.environ.get("RDB_BENCH_K", 10))
GEN_CHUNK = 250_000            # rows per generation chunk; chunk c uses seed DB_SEED + c on every rank
DB_SEED, Q_SEED, LABEL_SEED = 1234, 5678, 91011
METRIC_NAME = "QPS @k=10 on 10M x 768 DB (exact flat search)"
_CFG = "C3" if (N_DB, DIM, NQ, K) == (10_000_000, 768, 65536, 10) else \
    ("C5" if (N_DB, DIM, K) == (100_000_000, 256, 100) else "custom (RDB_BENCH_* overrides)")
WORKLOAD = f"{_CFG}: {N_DB}x{DIM} bf16 DB, {NQ}-query batch, k={K}, cosine (IP + L2-normalise), exact flat search"
if _CFG != "C3":
    METRIC_NAME = f"QPS @k={K} on {N_DB} x {DIM} DB (exact flat search)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this exact workload (profiles/r02_ncu_c3_traffic.json); None if it does not match."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_c3_traffic.json")
    try:
        with open(p) as f:
            j = json.load(f)
        if _CFG == "C3" and j.get("workload", "").startswith(f"C3: {N_DB}x{DIM} bf16 DB, {NQ}-query batch, k={K},"):
            return float(j["traffic_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"bf16_sustained": float(j.get("bf16_tflops_sustained", 1338.4)), "bf16_burst": float(j.get("bf16_tflops", 1634.0)),
                "hbm": float(j.get("hbm_gbs", 6547.8)), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs a few hundred ms to come up, longer on 8-GPU boxes) and only the
    samples whose timestamps fall inside [mark_begin, mark_end] are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.05]
        rows = inside if inside else [r for (_, r) in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "samples_inside_timed_region": len(inside), "reasons": sorted(reasons)}


def gen_db_chunk(torch, c, device):
    g = torch.Generator(device=device)
    g.manual_seed(DB_SEED + c)
    rows = min(GEN_CHUNK, N_DB - c * GEN_CHUNK)
    return torch.randn((rows, DIM), generator=g, device=device, dtype=torch.float32)


def gen_queries(torch, device):
    g = torch.Generator(device=device)
    g.manual_seed(Q_SEED)
    return torch.randn((NQ, DIM), generator=g, device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (faiss is not installable here, so: the oracle's
    torch-CPU restatement of faiss's blocked sgemm + top-k), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import numpy as np
    import torch
    orc = importlib.import_module("oracle.flat_oracle")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = min(N_DB, int(os.environ.get("RDB_REF_ROWS", 1_000_000)))
    nq_s = int(os.environ.get("RDB_REF_Q", 2048))
    rng = np.random.default_rng(DB_SEED)
    xb = rng.standard_normal((n_s, DIM), dtype=np.float32)
    xb /= (np.linalg.norm(xb, axis=1, keepdims=True) + 1e-12)
    xq = np.random.default_rng(Q_SEED).standard_normal((nq_s, DIM), dtype=np.float32)

    def step():
        qn = orc.maybe_normalize(xq, True)                       # vector_database.py:166
        return orc.torch_cpu_flat_search(xb, qn, K, orc.METRIC_IP)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    # flat search cost is linear in database rows: scale the sampled step to the full 10M-row database
    qps = nq_s / (dt * (N_DB / n_s))
    line = {"impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "faiss_present": False},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{nq_s} queries x {n_s} rows per step (fp32, torch-CPU sgemm+topk), "
                                       f"QPS scaled linearly in rows to {N_DB}"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def cpu_baseline(torch, orc, idx, q_dev):
    """Oracle port on the host cores on a bounded sample of the SAME workload (stored rows read back)."""
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = min(idx.ntotal, 1_000_000)
    ids = torch.arange(n_s, device=q_dev.device, dtype=torch.int64)
    xb = idx.reconstruct_batch(ids).cpu().numpy()                # stored (bf16-rounded, normalised) rows as fp32
    xq = q_dev[:16384].cpu().numpy()
    qn = orc.round_bf16(orc.maybe_normalize(xq, True))           # bf16 config: the oracle sees the rounded values
    orc.torch_cpu_flat_search(xb, qn[:64], K, orc.METRIC_IP)     # warm-up
    t0 = time.perf_counter()
    orc.torch_cpu_flat_search(xb, qn[:256], K, orc.METRIC_IP)    # calibration
    t_cal = time.perf_counter() - t0
    nq_s = int(max(256, min(len(qn), 256 * (12.0 / max(t_cal, 1e-3)))))   # ~12 s of CPU work
    t0 = time.perf_counter()
    Dc, Ic = orc.torch_cpu_flat_search(xb, qn[:nq_s], K, orc.METRIC_IP)
    dt = time.perf_counter() - t0
    qps = nq_s / (dt * (N_DB / n_s))
    return {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{nq_s} queries x {n_s} of {N_DB} rows in {dt:.2f} s (torch-CPU sgemm+topk, fp32), "
                      f"QPS scaled linearly in rows to the full database"}, (xb, qn[:nq_s], Ic)


def brute_force_topk(torch, idx, q_dev, nsub, row0, dist=None, world=1):
    """Checker: exact top-K of the first `nsub` queries by an fp32 matmul over the STORED rows of this rank's shard
    (read back through reconstruct_batch), merged over the ranks with one all-gather.  Returns (values, global ids)."""
    n = idx.ntotal
    dev = q_dev.device
    qs = torch.nn.functional.normalize(q_dev[:nsub], dim=1, eps=1e-12).to(torch.bfloat16).to(torch.float32)
    best_v = torch.full((nsub, K), float("-inf"), device=dev)
    best_i = torch.full((nsub, K), -1, dtype=torch.int64, device=dev)
    step = 250_000
    for s0 in range(0, n, step):
        e = min(n, s0 + step)
        rows = idx.reconstruct_batch(torch.arange(row0 + s0, row0 + e, device=dev, dtype=torch.int64))
        sc = qs @ rows.T
        v, i = torch.topk(sc, min(K, e - s0), dim=1)
        cv, ci = torch.cat([best_v, v], 1), torch.cat([best_i, i + row0 + s0], 1)
        best_v, sel = torch.topk(cv, K, dim=1)
        best_i = torch.gather(ci, 1, sel)
        del rows, sc
    if world > 1:
        gv = torch.empty((world,) + tuple(best_v.shape), device=dev)
        gi = torch.empty((world,) + tuple(best_i.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gv, best_v.contiguous())
        dist.all_gather_into_tensor(gi, best_i.contiguous())
        cv, ci = gv.permute(1, 0, 2).reshape(nsub, -1), gi.permute(1, 0, 2).reshape(nsub, -1)
        best_v, sel = torch.topk(cv, K, dim=1)
        best_i = torch.gather(ci, 1, sel)
    return best_v, best_i


def recall_at_k(torch, ours_i, ref_i):
    return (ours_i.unsqueeze(2) == ref_i.unsqueeze(1)).any(2).float().mean().item()


def _percentiles(ms):
    ms = sorted(ms)
    return ms[len(ms) // 2], ms[min(len(ms) - 1, int(len(ms) * 0.99))]


def single_process_e2e(torch, np, pkg, ngpus, q_np, steps, ref_ids):
    """C3 through ONE VectorDatabase whose index is row-sharded over all GPUs of the box by this one process
    (config.db_devices): pageable host numpy queries in, host numpy results out."""
    class _Cfg:
        vector_db_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"rdb_bench_sp_{os.getpid()}")
        vector_db_index_type = "IP"
        top_k = K
        db_dtype = "bf16"
        db_devices = list(range(ngpus))
    vdb = pkg.VectorDatabase(_Cfg())
    vdb.create_index(DIM)
    idx = vdb.index
    idx.reserve(N_DB)
    dev0 = torch.device("cuda", 0)
    for c in range(-(-N_DB // GEN_CHUNK)):
        idx.add(gen_db_chunk(torch, c, dev0), normalize=True)
    for _ in range(2):
        Dn, In = vdb.search_batch(q_np, k=K)
    t0 = time.perf_counter()
    for _ in range(steps):
        Dn, In = vdb.search_batch(q_np, k=K)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    out = {"value": NQ / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "n_gpus": ngpus,
           "shard_rows": idx.shard_sizes, "h2d_bytes_per_step": NQ * DIM * 4, "d2h_bytes_per_step": NQ * K * 16,
           "note": "one process, one VectorDatabase(db_devices=all GPUs), pageable numpy in/out: every GPU uploads its 1/N "
                   "slice of the batch from its own host thread, NVLink exchange, per-GPU search, per-GPU slice merge",
           "ids_equal_sharded_run": float((In[:len(ref_ids)] == ref_ids).all(1).mean()), "queries_compared": int(len(ref_ids))}
    vdb.cleanup_gpu_resources()
    return out


def secondary_block(torch, np, pkg, orc, dev, q_dev, pk):
    """BASELINE.json configs[1] (C2), configs[3] (C4) and north_star (a) (ingest) on this GPU, in this run."""
    out = {}
    n2, q2 = 1_000_000, 10_000
    vdbs = {}

    class _Cfg:
        def __init__(self, itype, dtype):
            self.vector_db_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"rdb_bench_sec_{os.getpid()}")
            self.vector_db_index_type, self.top_k, self.db_dtype = itype, K, dtype

    # ---- ingest (north_star a): 1M x 768 fp32 device rows -> bf16 store, with and without the fused normalisation
    x = torch.cat([gen_db_chunk(torch, c, dev) for c in range(n2 // GEN_CHUNK)], 0)
    for norm in (False, True):
        scratch = pkg.FlatIndex(DIM, pkg.METRIC_IP, "bf16", device=dev.index)
        scratch.reserve(4 * n2)
        scratch.add(x, normalize=norm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            scratch.add(x, normalize=norm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gbs = n2 * (DIM * 4 + DIM * 2 + 4) / (ms * 1e-3) / 1e9
        out["ingest_bf16_normalize" if norm else "ingest_bf16"] = {
            "rows": n2, "ms": ms, "GBs": gbs, "frac_of_hbm": gbs / pk["hbm"],
            "bytes_per_row": DIM * 4 + DIM * 2 + 4, "kernel": "ingest_rows_kernel (+ 2 norm-summary kernels)"}
        scratch.close()
    # ---- C2: 1M x 768 fp32 store, 10k-query batch, k=10, exact-fp32 neighbours (tiered certified tensor-core search)
    roof_c2 = pk["bf16_sustained"] * 1e12 / (2.0 * n2 * DIM)           # one-term tensor roofline, queries/s
    for name, metric, norm in (("c2_l2", pkg.METRIC_L2, False), ("c2_cosine", pkg.METRIC_IP, True)):
        idx = pkg.FlatIndex(DIM, metric, "f32", device=dev.index)
        idx.reserve(n2)
        idx.add(x, normalize=norm)
        q = q_dev[:q2]
        for _ in range(3):
            Dv, Iv = idx.search(q, K, normalize=norm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record()
        for _ in range(5):
            Dv, Iv = idx.search(q, K, normalize=norm)
            kms.append(idx.last_kernel_ms()[0])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        t1q, t1u = idx.last_tier1
        # checker: the oracle's CPU search over the stored rows for 128 of the queries
        xb = idx.reconstruct_batch(torch.arange(n2, device=dev, dtype=torch.int64)).cpu().numpy()
        qn = orc.maybe_normalize(q[:128].cpu().numpy(), norm)
        _, Ic = orc.torch_cpu_flat_search(xb, qn, K, metric)
        eq = float((np.asarray(Ic) == Iv[:128].cpu().numpy()).all(1).mean())
        out[name] = {"workload": f"C2: {n2}x{DIM} fp32 store, {q2}-query batch, k={K}, {'cosine' if norm else 'L2'}, exact fp32",
                     "qps": q2 / (ms * 1e-3), "ms_per_batch": ms, "scorer_kernel_ms": sum(kms) / len(kms),
                     "roofline_qps_one_term_tensor": roof_c2, "frac_of_roofline": q2 / (ms * 1e-3) / roof_c2,
                     "tier1_queries": t1q, "tier1_uncertified": t1u, "exact_fallback_queries": idx.last_uncertified,
                     "ids_equal_oracle_128q": eq}
        vdbs[name] = idx
        del xb
    # ---- C4: batch-1 latency, 10 000 sequential HOST queries through VectorDatabase.search (k = top_k + 10 = 15)
    bf = pkg.FlatIndex(DIM, pkg.METRIC_IP, "bf16", device=dev.index)
    bf.reserve(n2)
    bf.add(x, normalize=True)
    del x
    qh = np.array(q_dev[:10_000].cpu().numpy())                         # pageable host memory
    for name, idx, cos, bpe in (("c4_bf16", bf, True, 2), ("c4_fp32", vdbs["c2_l2"], False, 4)):
        vdb = pkg.VectorDatabase(_Cfg("IP" if cos else "L2", "bf16" if bpe == 2 else "f32"))
        vdb.index, vdb._cosine = idx, cos
        for i in range(200):
            vdb.search(qh[i], k=15)
        lat = []
        for i in range(len(qh)):
            t0 = time.perf_counter()
            vdb.search(qh[i], k=15)
            lat.append((time.perf_counter() - t0) * 1e3)
        p50, p99 = _percentiles(lat)
        kms = idx.last_kernel_ms()[0]
        floor = n2 * DIM * bpe / (pk["hbm"] * 1e9) * 1e3
        out[name] = {"workload": f"C4: {n2}x{DIM} {'bf16' if bpe == 2 else 'fp32'} store, 10000 sequential host queries, k=15",
                     "p50_ms": p50, "p99_ms": p99, "mean_ms": sum(lat) / len(lat), "kernel_ms_last": kms,
                     "hbm_floor_ms": floor, "p50_frac_of_floor": floor / p50, "kernel_frac_of_floor": floor / kms,
                     "qps_single_stream": 1e3 / p50}
        vdb.index = None
    # ---- mid-size batches (the reference's own batch is 256 segments, k_search = top_k + 10 = 15; pipeline.py:449-478):
    #      1M x 768 bf16, device in/out.  Floor = the slower of the tensor term at the sustained bf16 peak and the
    #      database bytes at the HBM copy peak.
    for qn_ in (128, 256, 512):
        q = q_dev[:qn_]
        for _ in range(5):
            bf.search(q, 15, normalize=True)
        torch.cuda.synchronize()
        kms, wall = [], []
        for _ in range(20):
            t0 = time.perf_counter()
            bf.search(q, 15, normalize=True)
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            kms.append(bf.last_kernel_ms()[0])
        kmed, wmed = sorted(kms)[len(kms) // 2], sorted(wall)[len(wall) // 2]
        floor = max(2.0 * qn_ * n2 * DIM / (pk["bf16_sustained"] * 1e12), n2 * DIM * 2 / (pk["hbm"] * 1e9)) * 1e3
        out[f"midbatch_q{qn_}"] = {"workload": f"{n2}x{DIM} bf16 store, {qn_}-query batch, k=15, cosine, device in/out",
                                   "search_ms_median": wmed, "scorer_kernel_ms_median": kmed, "floor_ms": floor,
                                   "kernel_frac_of_floor": floor / kmed, "search_frac_of_floor": floor / wmed,
                                   "qps": qn_ / (wmed * 1e-3)}
    for idx in list(vdbs.values()) + [bf]:
        idx.close()
    # ---- reference scale (the shapes the reference itself runs: 25 423 stored segments x 5376 features, batches of 256,
    #      top_k = 5 (+10 for self-exclusion); pipeline.py:449-532): retrieve_similar_vectors = search + exclusion filter +
    #      row gather + labels, device in/out, fp32 store (bit-exact rows) and bf16 store
    nr, dr, br = 25_423, 5376, 256
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    xr = torch.randn((nr, dr), generator=g, device=dev)
    qr = torch.randn((br, dr), generator=g, device=dev)
    qpaths = [f"/data/train/p{i}" for i in range(br)]          # the batch's own files are excluded from its neighbours
    for dtype in ("f32", "bf16"):
        vdb = pkg.VectorDatabase(_Cfg("L2", dtype))
        vdb.create_index(dr)
        vdb.add_vectors(xr, [f"p{i}" for i in range(nr)], [i & 1 for i in range(nr)], {})
        for _ in range(3):
            pkg.retrieve_similar_vectors(vdb, qr, 5, query_paths=qpaths)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            vec, lbl = pkg.retrieve_similar_vectors(vdb, qr, 5, query_paths=qpaths)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / 20
        out[f"refscale_retrieve_{dtype}"] = {
            "workload": f"retrieve_similar_vectors: {nr}x{dr} {dtype} store, {br} queries, top_k=5 (+10), L2, device in/out",
            "ms_per_batch": ms, "batches_per_s": 1e3 / ms, "search_kernel_ms": vdb.index.last_kernel_ms()[0],
            "neighbour_tensor_shape": list(vec.shape)}
        vdb.cleanup_gpu_resources()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C4 / ingest block (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)          # raises if libradad_flat.so is missing: no fallback
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's row shard on the device (never materialise the database on the host)
    sidx = pkg.ShardedFlatIndex(DIM, pkg.METRIC_IP, "bf16", device=local_rank,
                                exchange=os.environ.get("RDB_EXCHANGE", "peer"))
    start, end = sidx.set_shard(N_DB)
    lab_gen = torch.Generator(device=dev)
    for c in range(start // GEN_CHUNK, -(-end // GEN_CHUNK)):
        x = gen_db_chunk(torch, c, dev)
        lo, hi = max(start, c * GEN_CHUNK) - c * GEN_CHUNK, min(end, (c + 1) * GEN_CHUNK) - c * GEN_CHUNK
        sidx.add_local(x[lo:hi], normalize=True)                 # fused normalise + bf16 convert + |y|^2
        del x
    lab_gen.manual_seed(LABEL_SEED + rank)
    sidx.set_labels_local(torch.randint(0, 2, (end - start,), generator=lab_gen, device=dev).float().cpu().numpy())
    idx = sidx.local
    q_dev = gen_queries(torch, dev)
    # what a reference caller hands over: ordinary (pageable) host memory, never a pinned buffer
    q_host = q_dev.cpu()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sidx.search(q_dev, K, normalize=True)

    # the reference-facing call with HOST buffers
    if world == 1:
        class _Cfg:
            vector_db_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"rdb_bench_{os.getpid()}")
            vector_db_index_type = "IP"
            top_k = K
            db_dtype = "bf16"
        vdb = pkg.VectorDatabase(_Cfg())
        vdb.index = idx
        vdb._cosine = True
        q_np = q_host.numpy()

        def step_e2e():
            return vdb.search_batch(q_np, k=K)
    else:
        def step_e2e():
            # every rank uploads 1/G of the (pageable) host batch; one all-gather over NVLink assembles it on every GPU
            D, I, L = sidx.search_from_host(q_host, K, normalize=True)
            return D.cpu(), I.cpu()

    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    launches0 = idx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    barrier()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
        kern_ms.append(idx.last_kernel_ms()[0])                 # CUDA events around the scorer on its stream
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = idx.launch_count - launches0
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    _, algo, nsplits = idx.last_kernel_ms()

    # ---- end-to-end timing (host buffers)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oe = step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps

    # ---- correctness of THIS run's answer (every rank takes part: the brute force is sharded like the database)
    nchk = min(NQ, int(os.environ.get("RDB_BENCH_CHECK_Q", 4096)))
    corr = {"queries_checked": nchk}
    try:
        bv, bi = brute_force_topk(torch, idx, q_dev, nchk, start, dist if world > 1 else None, world)
        corr[f"recall_at_{K}"] = recall_at_k(torch, out[1][:nchk], bi)
        corr["ids_equal_bruteforce"] = float((out[1][:nchk] == bi).all(1).float().mean().item())
        corr["checker"] = ("fp32 matmul + top-k over the stored rows" +
                           (f", per-rank over the local shard, all-gather, merge ({world} ranks)" if world > 1 else ""))
        if world > 1:
            # the same queries on ONE GPU holding the whole database (rank 0 builds it; the other ranks wait)
            eq = torch.zeros((2,), device=dev)
            if rank == 0:
                full = pkg.FlatIndex(DIM, pkg.METRIC_IP, "bf16", device=local_rank)
                full.reserve(N_DB)
                for c in range(-(-N_DB // GEN_CHUNK)):
                    full.add(gen_db_chunk(torch, c, dev), normalize=True)
                D1, I1 = full.search(q_dev[:nchk], K, normalize=True)
                eq[0] = (I1 == out[1][:nchk]).all(1).float().mean()
                eq[1] = (D1 == out[0][:nchk]).all(1).float().mean()
                full.close()
            dist.broadcast(eq, 0)
            corr["ids_equal_n1_subset"] = float(eq[0].item())
            corr["distances_equal_n1_subset"] = float(eq[1].item())
    except Exception as e:  # noqa: BLE001
        corr["error"] = f"{type(e).__name__}: {e}"

    # ---- N > 1: the same workload through ONE process driving every GPU (how pipeline.py:90 constructs the database):
    #      rank 0 builds a second, row-sharded copy over all GPUs and times VectorDatabase.search_batch with pageable
    #      host numpy in / out; the other ranks wait on the CPU (TCP store), so no collective kernel occupies their GPUs
    sp = None
    if world > 1 and not args.no_secondary:
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                sp = single_process_e2e(torch, np, pkg, world, q_host.numpy(), args.steps, out[1][:4096].cpu().numpy())
            except Exception as e:  # noqa: BLE001
                sp = {"error": f"{type(e).__name__}: {e}"}
            store.set("rdb_bench_sp_done", "1")
        else:
            store.wait(["rdb_bench_sp_done"])

    t = torch.tensor([ms_dev, ms_e2e, sum(kern_ms) / len(kern_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_kern = [float(v) for v in t.tolist()]

    if rank == 0:
        pk = peaks()
        rows_local = end - start
        flops = 2.0 * NQ * rows_local * DIM
        ach = flops / (ms_kern * 1e-3) / 1e12
        line = {
            "metric": METRIC_NAME, "value": NQ / (ms_dev * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"row-shards x{world}", "scorer": algo,
                       "exchange": ("none" if world == 1 else sidx.exchange + ("-memory fused gather+merge kernel"
                                                                              if sidx.exchange == "peer" else " all-gather + merge")),
                       "db_chunks_per_query_tile": nsplits,
                       "l2_policy": "inputs larger than L2 (database shard >> 126 MB), no flush needed"},
            "e2e": {"value": NQ / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": NQ * DIM * 4, "d2h_bytes_per_step": NQ * K * 12 * world,
                    "note": ("pageable host numpy in/out through VectorDatabase.search_batch" if world == 1 else
                             "each rank uploads 1/N of the pageable host batch (total = h2d_bytes_per_step), NVLink "
                             "all-gather, sharded search, every rank reads the merged result back")},
            "gpu_launches": int(launches),
            "step_breakdown_ms": {"scorer_kernel": ms_kern, "rest_of_step": ms_dev - ms_kern,
                                  "note": "rest = query prep + candidate merge (+ at N > 1: barrier all-reduce, per-rank slice "
                                          "merge over peer memory, all-gather of the result slices); max over ranks"},
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["bf16_sustained"], "traffic": ncu_traffic() if world == 1 else None,
                         "traffic_note": "dram read+write bytes/launch from the committed ncu --set full capture of this "
                                         "workload (profiles/r02_ncu_c3_traffic.json): 81.3 GB read = 5.3 database volumes "
                                         "(each 15.36 GB chunk pass is shared by the 74 CTA pairs of a scheduling slot; 256 "
                                         "query-tile pairs x 13 chunks touch 4.5 slots per chunk), 0.30 GB written (the "
                                         "candidate lists); ~2 % of DRAM bandwidth, not the bound",
                         "kernel": "score_select_tc_kernel", "kernel_ms": ms_kern,
                         "flops_per_launch": flops, "peak_source": pk["src"] + " sustained bf16",
                         "frac_of_burst": ach / pk["bf16_burst"],
                         "hbm_frac": (rows_local * DIM * 2 / (ms_kern * 1e-3) / 1e9) / pk["hbm"]},
        }
        line["correctness"] = corr
        if sp is not None:
            line["e2e_single_process"] = sp
        if world == 1:
            orc = importlib.import_module("oracle.flat_oracle")
            if not args.no_cpu_baseline:
                cb, (xb_s, qn_s, Ic) = cpu_baseline(torch, orc, idx, q_dev)
                # the oracle's answer is a checker too: the CUDA path over the SAME rows (the first 1M stored rows in an
                # index of their own) must return the oracle's neighbours for the sampled queries
                sub = pkg.FlatIndex(DIM, pkg.METRIC_IP, "bf16", device=local_rank)
                sub.add(torch.from_numpy(xb_s).to(dev))
                _, Is = sub.search(torch.from_numpy(np.ascontiguousarray(qn_s)).to(dev), K)
                Is = Is.cpu().numpy()
                Ic = np.asarray(Ic)
                cb["ids_equal_cuda_vs_oracle"] = float((Is == Ic).all(1).mean())
                cb["recall_cuda_vs_oracle"] = float(np.mean([len(set(a) & set(b)) / K for a, b in zip(Is, Ic)]))
                sub.close()
                line["cpu_baseline"] = cb
            if not args.no_secondary:
                try:
                    line["secondary"] = secondary_block(torch, np, pkg, orc, dev, q_dev, pk)
                except Exception as e:  # noqa: BLE001
                    line["secondary_error"] = f"{type(e).__name__}: {e}"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
