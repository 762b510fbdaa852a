"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden vectors and the oracle.

Tolerances (BASELINE.json north_star): neighbour ids identical except for ties within 1e-5 relative distance
for fp32 storage / 1e-3 for bf16 storage; integer-lattice inputs (exact arithmetic) must match BIT-EXACTLY,
ties broken by the lowest id.
"""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, Cfg

pytestmark = pytest.mark.gpu
SEARCH = sorted(glob.glob(os.path.join(GOLDEN, "search_*.npz")))
TOL_F32, TOL_BF16 = 1e-5, 1e-3


def _load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def _fill(vdb, g):
    n = g["xb"].shape[0]
    vdb.add_vectors(g["xb"], [f"/data/spk{i % 13}/utt_{i:05d}.wav" for i in range(n)],
                    [int(v) for v in g["labels"]], {"speaker_id": [f"spk{i % 13}" for i in range(n)]})


@pytest.mark.parametrize("path", SEARCH, ids=[os.path.basename(p)[7:-4] for p in SEARCH])
def test_golden_fp32(pkg, oracle, path, tmp_path):
    """fp32 store: the drop-in VectorDatabase reproduces what the reference wrapper returned."""
    g = _load(path)
    cfg = Cfg(tmp_path / "g", str(g["index_type"]), normalize_for_ip=bool(g["normalize_for_ip"]),
              vector_add_batch_size=256)
    vdb = pkg.VectorDatabase(cfg)
    _fill(vdb, g)
    k = int(g["k"])
    D, I = vdb.search_batch(g["xq"], k=k)
    assert D.dtype == np.float32 and I.dtype == np.int64
    assert D.shape == g["dist"].shape and I.shape == g["idx"].shape
    assert vdb.index.ntotal == int(g["ntotal"]) and bool(vdb._cosine) == bool(g["cosine"])
    assert len(vdb.vector_paths) == int(g["n_paths"])
    name = os.path.basename(path)
    metric = oracle.METRIC_IP if str(g["index_type"]) == "IP" else oracle.METRIC_L2
    if "lattice" in name or "kat_tiny" in name:
        np.testing.assert_array_equal(I, g["idx"])
        np.testing.assert_array_equal(D, g["dist"])
    else:
        ref = oracle.FlatIndexOracle(g["xb"].shape[1], metric)
        ref.add(oracle.maybe_normalize(g["xb"], bool(g["cosine"])))
        qn = oracle.maybe_normalize(g["xq"], bool(g["cosine"]))
        st = oracle.compare_topk(D, I, g["dist"], g["idx"], lambda ids: ref.exact_scores(qn, ids), metric,
                                 tol=TOL_F32, abs_floor=2e-5 if metric == oracle.METRIC_L2 else 1e-6)
        assert st["recall"] == 1.0
    d1, i1 = vdb.search(g["xq"][0], k=k)
    assert d1.shape == g["dist_single"].shape
    np.testing.assert_array_equal(i1, g["idx_single"])
    dd, di = vdb.search_batch(g["xq"][:2])
    np.testing.assert_array_equal(di, g["idx_default"])
    rec = np.stack([vdb.index.reconstruct(int(i)) for i in I[0]])
    np.testing.assert_allclose(rec, g["recon_row0"], rtol=1e-6, atol=1e-7)
    # labels gathered on the device == labels[idx]
    D2, I2, L2 = vdb.search_batch_with_labels(g["xq"], k=k)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(L2, g["labels"][I].astype(np.float32))


@pytest.mark.parametrize("store", ["bf16", "f16"])
@pytest.mark.parametrize("algo", ["simt", "tc"])
@pytest.mark.parametrize("name", ["lattice_l2", "lattice_ip"])
def test_lattice_bit_exact_16bit_stores(pkg, name, algo, store):
    """{-2..2} lattice: every product and sum is exact in bf16/f16 x fp32 -> tensor-core and CUDA-core scorers
    must return the golden ids and distances bit-for-bit (duplicate rows exercise the lowest-id tie rule)."""
    g = _load(os.path.join(GOLDEN, f"search_{name}.npz"))
    metric = pkg.METRIC_IP if name.endswith("ip") else pkg.METRIC_L2
    idx = pkg.FlatIndex(g["xb"].shape[1], metric, store)
    idx.add(g["xb"][:300])
    idx.add(g["xb"][300:])
    D, I = idx.search(g["xq"], int(g["k"]), algo=algo)
    np.testing.assert_array_equal(I, g["idx"])
    np.testing.assert_array_equal(D, g["dist"])


def _gauss(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


CASES = [
    # name            N      D     Q    k   metric cos    store   algo
    ("c1_cos_f32",    20000, 768,  1000, 10, "IP", True,  "f32",  "simt"),   # BASELINE configs[0]
    ("l2_f32_ragged", 5003,  200,  77,   15, "L2", False, "f32",  "simt"),
    ("ip_f32_oddD",   3001,  101,  33,   7,  "IP", False, "f32",  "simt"),   # D % 4 != 0
    ("l2_bf16_tc",    50000, 768,  512,  10, "L2", False, "bf16", "tc"),
    ("cos_bf16_tc",   50000, 768,  300,  10, "IP", True,  "bf16", "tc"),
    ("l2_bf16_tc_k32", 9000, 256,  130,  32, "L2", False, "bf16", "tc"),
    ("ip_bf16_tc_k1", 4097,  64,   129,  1,  "IP", False, "bf16", "tc"),
    ("l2_bf16_oddD",  6000,  100,  64,   15, "L2", False, "bf16", "tc"),     # D % 8 != 0 -> padded pitch
    ("l2_f16_tc",     8000,  320,  96,   16, "L2", False, "f16",  "tc"),
    ("refD_bf16_tc",  2048,  5376, 256,  15, "L2", False, "bf16", "tc"),     # reference D = 7*768, Q, k
    ("refD_f32",      1500,  3584, 40,   15, "L2", False, "f32",  "simt"),   # Whisper D = 7*512
    ("l2_bf16_simt",  7000,  192,  70,   17, "L2", False, "bf16", "simt"),
    # fp32 store on the tensor cores: split-precision (hi/lo bf16, 3 MMAs) + exact fp32 re-rank + certificate
    ("cos_f32_split", 60000, 768,  300,  10, "IP", True,  "f32",  "tc"),     # C2-like, kc = 16
    ("l2_f32_split",  30000, 768,  200,  15, "L2", False, "f32",  "tc"),     # reference k = top_k + 10, kc = 32
    ("l2_f32_split_ragged", 5003, 200, 77, 24, "L2", False, "f32", "tc"),
    ("ip_f32_split_q1", 20000, 256, 1,   5,  "IP", False, "f32",  "tc"),
    ("l2_f32_auto",   20000, 768,  1000, 10, "L2", False, "f32",  "auto"),
    ("l2_f32_split_k40",  30000, 256, 200, 40,  "L2", False, "f32", "tc"),   # kc = 64 (reservoir epilogue + re-rank of 64)
    ("cos_f32_split_k100", 40000, 128, 150, 100, "IP", True, "f32", "tc"),   # kc = 128
    ("ip_bf16_simt_k100", 5000, 128, 50, 100, "IP", False, "bf16", "simt"),
    # small-batch HBM-streaming scorer (batch-1 latency path; exact fp32 for fp32 stores)
    ("stream_l2_f32_q1",   30000, 768, 1, 15, "L2", False, "f32",  "stream"),
    ("stream_cos_f32_q2",  30000, 768, 2, 15, "IP", True,  "f32",  "stream"),
    ("stream_l2_bf16_q1",  50001, 768, 1, 15, "L2", False, "bf16", "stream"),
    ("stream_ip_bf16_q3",  7003,  100, 3, 32, "IP", False, "bf16", "stream"),   # Dp = 104, ragged N
    ("stream_l2_f16_q4",   9000,  256, 4, 5,  "L2", False, "f16",  "stream"),
    ("auto_l2_f32_q1",     30000, 768, 1, 15, "L2", False, "f32",  "auto"),
    ("stream_ip_bf16_q1_k100", 40000, 256, 1, 100, "IP", True, "bf16", "stream"),   # C5, Q = 1
    ("stream_l2_f32_q4_k128",  9000,  128, 4, 128, "L2", False, "f32", "stream"),
    ("stream_l2_bf16_q2_k33",  5001,  64,  2, 33,  "L2", False, "bf16", "stream"),
    # large k on a large database: sampled pivot + filter pass (N >= 131072, k > 32)
    # (the pivot pass samples every m-th warp step, m = min(64, steps per warp), and needs 16 m >= 4 k: millions of rows)
    ("stream_filter_cos_bf16_q1_k100", 2000000, 64, 1, 100, "IP", True,  "bf16", "stream"),
    ("stream_filter_l2_f32_q3_k64",    1000000, 64, 3, 64,  "L2", False, "f32",  "stream"),
    ("stream_list4_l2_f16_q4_k128",    140001, 72,  4, 128, "L2", False, "f16",  "stream"),   # too few rows: LIST policy
    # large k on the tensor cores: local-memory reservoir + exact bisection prune (C5: k = 100, D = 256)
    ("ip_bf16_tc_k100", 60000, 256, 300, 100, "IP", True,  "bf16", "tc"),
    ("l2_bf16_tc_k64",  20000, 128, 130, 64,  "L2", False, "bf16", "tc"),
    ("l2_bf16_tc_k128", 9000,  96,  40,  128, "L2", False, "bf16", "tc"),
    ("ip_f16_tc_k33",   5000,  64,  33,  33,  "IP", False, "f16",  "tc"),
    # large k on a large shard: admission bound seeded from a strided 1/64 tile sample (N >= 262144)
    ("ip_bf16_tc_k100_pivot", 300000, 64, 200, 100, "IP", True,  "bf16", "tc"),
    ("l2_bf16_tc_k64_pivot",  280003, 40, 140, 64,  "L2", False, "bf16", "tc"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_seeded_gaussian_vs_oracle(pkg, oracle, case):
    name, N, Dm, Q, k, metric_s, cos, store, algo = case
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    xb, xq = _gauss(N, Dm, 1234), _gauss(Q, Dm, 5678)
    xq[::7] = xb[: len(xq[::7])] + 0.05 * _gauss(len(xq[::7]), Dm, 9)      # near-duplicates of DB rows
    xq[min(1, Q - 1)] = xb[5]                                                # exact self-match (distance 0)
    idx = pkg.FlatIndex(Dm, metric, store)
    for s in range(0, N, 4096):
        idx.add(xb[s:s + 4096], normalize=cos)
    D, I = idx.search(xq, k, normalize=cos, algo=algo)
    # oracle sees the same stored values: normalise in fp32, then the store's rounding
    ref = oracle.FlatIndexOracle(Dm, metric, store=store)
    ref.add(oracle.maybe_normalize(xb, cos))
    qn = oracle.maybe_normalize(xq, cos)
    Dr, Ir = ref.search(qn, min(k + 8, N), direct=False)
    tol = TOL_F32 if store == "f32" else TOL_BF16
    # L2 is evaluated as |q|^2 + |y|^2 - 2 q.y (as faiss does): near-duplicates cancel catastrophically, so the
    # absolute error floor scales with the magnitude of the cancelled terms -- a few fp32 ulps for the exact path,
    # ~2^-15 for the tensor cores (their fp32 accumulation truncates: measured 2e-5 at D = 5376).
    if metric == pkg.METRIC_L2:
        scale = float((qn * qn).sum(1).max() + (ref._base() ** 2).sum(1).max())
        floor = (2e-6 if (store == "f32" or algo == "simt") else 1e-4) * scale
    else:
        floor = 1e-6
    st = oracle.compare_topk(D, I, Dr, Ir, lambda ids: ref.exact_scores(qn, ids), metric, tol=tol, abs_floor=floor)
    assert st["recall"] >= 0.999, st
    # the stored rows come back as the oracle's rounded rows
    rec = idx.reconstruct_batch(I[0])
    np.testing.assert_allclose(rec, ref.reconstruct_batch(I[0]), rtol=2e-6 if store == "f32" else 1e-2, atol=1e-7)


@pytest.mark.parametrize("store", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("name", ["lattice_l2", "lattice_ip"])
def test_lattice_bit_exact_stream(pkg, name, store):
    """Streaming scorer on lattice data, 1..4 queries per call: bit-exact ids and distances, lowest id on ties."""
    g = _load(os.path.join(GOLDEN, f"search_{name}.npz"))
    metric = pkg.METRIC_IP if name.endswith("ip") else pkg.METRIC_L2
    idx = pkg.FlatIndex(g["xb"].shape[1], metric, store)
    idx.add(g["xb"])
    k = int(g["k"])
    for q0, nq in ((0, 1), (1, 2), (3, 3), (6, 4)):
        D, I = idx.search(g["xq"][q0:q0 + nq], k, algo="stream")
        np.testing.assert_array_equal(I, g["idx"][q0:q0 + nq])
        np.testing.assert_array_equal(D, g["dist"][q0:q0 + nq])


@pytest.mark.parametrize("metric_s", ["L2", "IP"])
def test_tc_pivot_fallback_when_sample_overestimates(pkg, oracle, metric_s):
    """Large-k tensor-core path with the sampled admission bound: 20 copies of a query sit in the FIRST (sampled) tile,
    so the sample's 16th best key is the exact-match score and only 20 < k rows of the whole shard reach it.  The
    device-side completeness check must then rerun the batch without the bound; the other queries (heavy lattice ties)
    take the normal path.  Ids and distances must equal the oracle bit-for-bit either way."""
    rng = np.random.default_rng(5)
    N, Dm, Q, k = 270000, 32, 150, 50
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(Q, Dm)).astype(np.float32)
    special = np.full((Dm,), 2.0, dtype=np.float32)
    special[::2] = -2.0
    xb[:20] = special
    xq[7] = special
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(Dm, metric, "bf16")
    idx.add(xb)
    ref = oracle.FlatIndexOracle(Dm, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    for _ in range(2):
        D, I = idx.search(xq, k, algo="tc")
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D, Dr)
    np.testing.assert_array_equal(Ir[7, :20], np.arange(20))
    D2, I2 = idx.search(xq[8:140], k, algo="tc")                  # no special query: the bound holds, no rerun
    np.testing.assert_array_equal(I2, Ir[8:140])
    np.testing.assert_array_equal(D2, Dr[8:140])


@pytest.mark.parametrize("store", ["bf16", "f32"])
@pytest.mark.parametrize("metric_s", ["L2", "IP"])
def test_stream_filter_fallback_on_heavy_ties(pkg, oracle, metric_s, store):
    """Streaming scorer, large k, large N, lattice data: only a handful of distinct scores exist, so far more than the
    4096 candidate slots tie with the sampled pivot -> the filter pass must raise its fallback flag and the list pass
    must deliver the exact result (ids and distances bit-for-bit, lowest id on ties).  A second query set whose
    candidates fit exercises the filter path itself, and repeating the calls checks the self-resetting counters."""
    rng = np.random.default_rng(21)
    N, Dm, k = 2000000, 16, 50
    xb = rng.integers(-1, 2, size=(N, Dm)).astype(np.float32)
    xq = rng.integers(-1, 2, size=(3, Dm)).astype(np.float32)
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(Dm, metric, store)
    idx.add(xb)
    ref = oracle.FlatIndexOracle(Dm, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    for _ in range(2):
        D, I = idx.search(xq, k, algo="stream")
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D, Dr)
    # distinct scores: every row gets its own tiny integer offset in one extra-wide coordinate -> no ties at all
    extra = np.zeros((N, 4), dtype=np.float32)
    extra[:, 0] = np.arange(N, dtype=np.float32) % 4093
    xb2 = np.concatenate([xb, extra], axis=1)
    xq2 = np.concatenate([xq, np.ones((3, 4), dtype=np.float32)], axis=1)
    idx2 = pkg.FlatIndex(Dm + 4, pkg.METRIC_IP, "f32")
    idx2.add(xb2)
    ref2 = oracle.FlatIndexOracle(Dm + 4, pkg.METRIC_IP)
    ref2.add(xb2)
    Dr2, Ir2 = ref2.search(xq2, k, direct=False)
    for nq in (3, 1, 2):
        D2, I2 = idx2.search(xq2[:nq], k, algo="stream")
        np.testing.assert_array_equal(I2, Ir2[:nq])
        np.testing.assert_array_equal(D2, Dr2[:nq])


@pytest.mark.parametrize("metric_s", ["L2", "IP"])
@pytest.mark.parametrize("k", [40, 100])
def test_lattice_bit_exact_large_k(pkg, oracle, metric_s, k):
    """Large-k tensor-core path on lattice data: heavy ties (few distinct scores) stress the reservoir prune's
    'earliest arrival wins' rule; result must equal the oracle bit-for-bit (ids and distances)."""
    rng = np.random.default_rng(7)
    xb = rng.integers(-2, 3, size=(6000, 64)).astype(np.float32)
    xb[3000:3200] = xb[10]                                    # 200 identical rows: ties far beyond k
    xq = rng.integers(-2, 3, size=(70, 64)).astype(np.float32)
    xq[0] = xb[10]
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(64, metric, "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k, algo="tc")
    ref = oracle.FlatIndexOracle(64, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)


@pytest.mark.parametrize("store", ["bf16", "f32"])
@pytest.mark.parametrize("metric_s", ["L2", "IP"])
@pytest.mark.parametrize("k", [10, 24, 100])
def test_cta_pair_kernel_bit_exact(pkg, oracle, metric_s, k, store):
    """The cta_group::2 form of the tensor-core scorer (a CTA pair runs M = 256 MMAs, RDB_TC_CG=2): lattice data, ragged
    query count (an odd number of 128-query tiles, last tile partly empty) and ragged N -- ids and distances must equal
    the oracle bit-for-bit, and the single-CTA form must agree."""
    if store == "f32" and k > 104:
        pytest.skip("split-precision path serves k <= 104")
    rng = np.random.default_rng(11)
    N, Dm, nq = 7013, 200, 300
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    xb[4000:4150] = xb[5]
    xq = rng.integers(-2, 3, size=(nq, Dm)).astype(np.float32)
    xq[299] = xb[5]
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(Dm, metric, store)
    idx.add(xb)
    ref = oracle.FlatIndexOracle(Dm, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    idx.set_option("tc_cta_group", 2)
    D2, I2 = idx.search(xq, k, algo="tc")
    idx.set_option("tc_cta_group", 1)
    D1, I1 = idx.search(xq, k, algo="tc")
    np.testing.assert_array_equal(I2, Ir)
    np.testing.assert_array_equal(D2, Dr)
    np.testing.assert_array_equal(I1, Ir)
    np.testing.assert_array_equal(D1, Dr)


@pytest.mark.parametrize("name", ["lattice_l2", "lattice_ip"])
def test_lattice_bit_exact_f32_split(pkg, name):
    """fp32 store, tensor-core split path: lattice values have lo == 0, all three MMA terms are exact."""
    g = _load(os.path.join(GOLDEN, f"search_{name}.npz"))
    metric = pkg.METRIC_IP if name.endswith("ip") else pkg.METRIC_L2
    idx = pkg.FlatIndex(g["xb"].shape[1], metric, "f32")
    idx.add(g["xb"])
    D, I = idx.search(g["xq"], int(g["k"]), algo="tc")
    np.testing.assert_array_equal(I, g["idx"])
    np.testing.assert_array_equal(D, g["dist"])


@pytest.mark.parametrize("k", [30, 64, 104])
@pytest.mark.parametrize("name", ["L2", "IP"])
def test_lattice_bit_exact_f32_split_large_k(pkg, oracle, name, k):
    """fp32 store, certified split path with kc = 64 / 128 candidates: lattice data (heavy ties -> most queries cannot
    be certified and take the exact fallback, the rest stay on the tensor cores) must match the oracle bit-for-bit."""
    rng = np.random.default_rng(31)
    N, Dm, nq = 6000, 48, 140
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(nq, Dm)).astype(np.float32)
    metric = pkg.METRIC_IP if name == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(Dm, metric, "f32")
    idx.add(xb)
    ref = oracle.FlatIndexOracle(Dm, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    D, I = idx.search(xq, k, algo="tc")
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)


def test_split_path_certificate_fallback(pkg):
    """A cluster of > kc identical rows cannot be certified (the kc-th candidate ties the k-th): those queries must
    take the exact fallback and still return the lowest ids; well-separated queries stay on the fast path."""
    N, Dm, k = 8000, 128, 10
    xb = _gauss(N, Dm, 1)
    xb[100:150] = xb[100]                      # 50 identical rows
    xq = _gauss(64, Dm, 2)
    xq[3] = xb[100]
    xq[9] = xb[100] + 1e-4
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "f32")
    idx.add(xb)
    Dt, It = idx.search(xq, k, algo="tc")
    unc = idx.last_uncertified
    Ds, Is = idx.search(xq, k, algo="simt")
    assert 2 <= unc <= 8, unc
    np.testing.assert_array_equal(It[3], np.arange(100, 110))
    np.testing.assert_array_equal(It[[3, 9]], Is[[3, 9]])
    np.testing.assert_array_equal(Dt[[3, 9]], Ds[[3, 9]])
    np.testing.assert_array_equal(It, Is)
    np.testing.assert_allclose(Dt, Ds, rtol=1e-5, atol=1e-4)


def test_wrapper_behaviour(pkg, tmp_path):
    with open(os.path.join(GOLDEN, "wrapper_behaviour.json")) as f:
        beh = json.load(f)
    cfg = Cfg(tmp_path / "w", "L2")
    vdb = pkg.VectorDatabase(cfg)
    assert vdb.index is None and vdb.gpu_index is None and vdb.vector_paths == [] and vdb.vector_metadata == {}
    assert vdb.db_path.endswith("faiss_index.bin") and vdb.metadata_path.endswith("metadata.pkl")
    with pytest.raises(ValueError) as e:
        vdb.search_batch(np.zeros((1, 8), np.float32))
    assert str(e.value) == beh["empty_search_error"]
    vdb.add_vectors(np.zeros((0, 8), np.float32), [], [], {})
    assert (vdb.index is None) == beh["index_none_after_empty_add"]
    xb = _gauss(10, 8, 1)
    vdb.add_vectors_batch(xb, [f"p{i}" for i in range(10)], list(range(10)),
                          {"split": 7, "speaker_id": [f"s{i}" for i in range(10)]}, batch_size=4)
    assert vdb.vector_metadata["split"] == beh["meta_split"]
    assert vdb.vector_metadata["speaker_id"] == beh["meta_speaker"]
    assert vdb.vector_labels == beh["labels"] and vdb.index.ntotal == 10 and vdb.index.d == 8
    d0, i0 = vdb.search_batch(xb[:2], k=0)
    assert [list(d0.shape), list(i0.shape)] == beh["k0_shapes"]
    assert [str(d0.dtype), str(i0.dtype)] == beh["k0_dtypes"]
    dk, ik = vdb.search_batch(xb[:3], k=50)                       # k clamped to ntotal (:169)
    assert dk.shape == (3, 10) and sorted(ik[0].tolist()) == list(range(10))
    assert ik[0, 0] == 0 and dk[0, 0] == 0.0
    vdb.save()
    with open(vdb.metadata_path, "rb") as f:
        import pickle
        assert sorted(pickle.load(f).keys()) == beh["pickle_keys"]
    v2 = pkg.VectorDatabase(cfg)
    v2.load()
    assert v2.index.ntotal == beh["loaded_ntotal"] and hasattr(v2, "_cosine") == beh["loaded_has_cosine_attr"]
    assert v2.vector_labels == beh["loaded_labels"]
    d2, i2 = v2.search_batch(xb[:3], k=4)
    d1, i1 = vdb.search_batch(xb[:3], k=4)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    np.testing.assert_array_equal(v2.index.reconstruct(3), xb[3])
    with pytest.raises(ValueError) as e:
        pkg.VectorDatabase(Cfg(tmp_path / "b", "HNSW")).create_index(8)
    assert str(e.value) == beh["bad_type_error"]
    mem = vdb.get_gpu_memory_usage()
    assert set(mem) == {"used", "total", "utilization"} and 0 < mem["utilization"] < 1
    # a failing slice is logged and skipped, lists not extended (:147-149)
    vdb.add_vectors_batch(_gauss(3, 9, 2), ["a", "b", "c"], [1, 1, 1], {})
    assert vdb.index.ntotal == 10 and len(vdb.vector_paths) == 10
    v3 = pkg.VectorDatabase(Cfg(tmp_path / "nothing", "L2"))
    v3.load()                                                     # missing files: warn, never raise
    assert v3.index is None


def test_cosine_after_load_quirk(pkg, tmp_path):
    g = _load(os.path.join(GOLDEN, "quirk_cosine_after_load.npz"))
    cfg = Cfg(tmp_path / "q", "IP")
    vdb = pkg.VectorDatabase(cfg)
    vdb.add_vectors(g["xb"], [f"p{i}" for i in range(200)], [0] * 200, {})
    d_b, i_b = vdb.search_batch(g["xq"], k=5)
    vdb.save()
    v2 = pkg.VectorDatabase(cfg)
    v2.load()
    d_a, i_a = v2.search_batch(g["xq"], k=5)
    np.testing.assert_array_equal(i_b, g["i_before"])
    np.testing.assert_array_equal(i_a, g["i_after"])
    np.testing.assert_allclose(d_b, g["d_before"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d_a, g["d_after"], rtol=1e-5, atol=1e-6)
    cfg.restore_cosine_on_load = True                             # opt-in fix
    v3 = pkg.VectorDatabase(cfg)
    v3.load()
    d_f, i_f = v3.search_batch(g["xq"], k=5)
    np.testing.assert_array_equal(i_f, g["i_before"])


def test_faiss_file_layout(pkg, tmp_path):
    """faiss IndexFlat layout: fourcc, d, ntotal, 2 dummies, is_trained, metric_type, count, fp32 rows."""
    import struct
    xb = _gauss(37, 12, 4)
    idx = pkg.FlatIndex(12, pkg.METRIC_IP, "f32")
    idx.add(xb)
    p = str(tmp_path / "faiss_index.bin")
    idx.save(p)
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxFI"
    d, n, _, _, trained, metric, count = struct.unpack_from("<iqqqBiQ", raw, 4)
    assert (d, n, trained, metric, count) == (12, 37, 1, 0, 37 * 12)
    body = np.frombuffer(raw, dtype=np.float32, offset=4 + 4 + 8 + 8 + 8 + 1 + 4 + 8)
    np.testing.assert_array_equal(body.reshape(37, 12), xb)
    back = pkg.FlatIndex.load(p, "bf16")
    assert back.ntotal == 37 and back.d == 12 and back.metric == pkg.METRIC_IP and back.store == "bf16"


@pytest.mark.parametrize("name", ["retrieve_l2", "retrieve_cos"])
def test_retrieve_similar_vectors_matches_reference_caller(pkg, name, tmp_path):
    """Our device-resident retrieve_similar_vectors vs what the reference's pipeline.py:449-532 returned."""
    import torch
    g = _load(os.path.join(GOLDEN, f"{name}.npz"))
    K, D = int(g["K"]), g["xb"].shape[1]
    cfg = Cfg(tmp_path / "r", str(g["index_type"]), top_k=K)
    vdb = pkg.VectorDatabase(cfg)
    paths = [str(p) for p in g["paths"]]
    labels = [torch.tensor(int(l)) for l in g["labels"]]           # 0-d tensors, as pipeline.py:436-441 stores
    vdb.add_vectors(g["xb"], paths, labels, {"speaker_id": ["s"] * len(paths)})
    qpaths = [str(p) for p in g["qpaths"]]
    train_ids = {str(s) for s in g["train_ids"]}
    q = torch.from_numpy(g["q"]).cuda()
    for tag, kw in (("excl_paths", dict(query_paths=qpaths, exclude_self=True)),
                    ("excl_train", dict(query_paths=None, exclude_self=True, training_file_ids=train_ids)),
                    ("noexcl", dict(query_paths=qpaths, exclude_self=False))):
        vec, lbl, pth, dst = pkg.retrieve_similar_vectors(vdb, q, K, D, return_info=True, return_distances=True, **kw)
        assert vec.is_cuda and vec.dtype == torch.float32 and tuple(vec.shape) == (len(qpaths), K, D)
        assert [list(r) for r in pth] == [list(map(str, r)) for r in g[f"{tag}_paths"]]
        np.testing.assert_array_equal(lbl.cpu().numpy(), g[f"{tag}_lbl"])
        # fp32 store: bit-identical neighbour tensors => identical RADADModel logits downstream, checked with the seeded
        # reference model itself (torch.jit trace written by make_golden.py)
        np.testing.assert_array_equal(vec.cpu().numpy(), g[f"{tag}_vec"])
        if tag == "excl_paths":
            model = torch.jit.load(os.path.join(GOLDEN, f"radad_model_{name}.pt")).eval()
            with torch.no_grad():
                logits = model(vec.cpu(), torch.from_numpy(g["q"]))
            np.testing.assert_array_equal(logits.numpy(), g["logits_excl_paths"])
        np.testing.assert_allclose(dst.cpu().numpy(), g[f"{tag}_dist"], rtol=1e-4, atol=2e-4, equal_nan=True)
    # return arities (pipeline.py:526-532)
    assert len(pkg.retrieve_similar_vectors(vdb, q, K, D, query_paths=qpaths)) == 2
    assert len(pkg.retrieve_similar_vectors(vdb, q, K, D, query_paths=qpaths, return_info=True)) == 3
    assert len(pkg.retrieve_similar_vectors(vdb, q, K, D, query_paths=qpaths, return_distances=True)) == 3
    empty = pkg.VectorDatabase(Cfg(tmp_path / "e", "L2", top_k=K))
    v, l, p, d = pkg.retrieve_similar_vectors(empty, q, K, D, return_info=True, return_distances=True)
    assert float(v.abs().sum()) == 0 and p[0] == [""] * K and bool(torch.isnan(d).all())


def test_device_tensor_path_and_vote(pkg):
    import torch
    xb, xq = _gauss(3000, 64, 1), _gauss(40, 64, 2)
    labels = (np.arange(3000) % 3 == 0).astype(np.float32)
    a = pkg.FlatIndex(64, pkg.METRIC_L2, "bf16")
    a.add(xb)
    a.set_labels(labels)
    b = pkg.FlatIndex(64, pkg.METRIC_L2, "bf16")
    b.add(torch.from_numpy(xb).cuda())                             # ingest straight from a CUDA tensor (f4)
    b.set_labels(labels)
    Dh, Ih, Lh = a.search(xq, 10, return_labels=True)
    Dd, Id, Ld = b.search(torch.from_numpy(xq).cuda(), 10, return_labels=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(Id.cpu().numpy(), Ih)
    np.testing.assert_array_equal(Dd.cpu().numpy(), Dh)
    np.testing.assert_array_equal(Lh, labels[Ih])
    vote = b.label_vote(Ld, 5)
    np.testing.assert_allclose(vote.cpu().numpy(), labels[Ih][:, :5].sum(1))


def test_sharded_single_process_equals_unsharded(pkg):
    """Row shards + merge kernel == unsharded search (G = 4 shards emulated on one GPU, no collective)."""
    import torch
    N, Dm, Q, k = 10007, 128, 65, 10
    xb, xq = _gauss(N, Dm, 1), _gauss(Q, Dm, 2)
    xb[5000] = xb[3]
    xq[0] = xb[3]                                                  # tie across shards -> lowest global id first
    labels = (np.arange(N) % 2).astype(np.float32)
    full = pkg.FlatIndex(Dm, pkg.METRIC_L2, "bf16")
    full.add(xb)
    full.set_labels(labels)
    Df, If, Lf = full.search(xq, k, return_labels=True)
    G = 4
    keys, gids, labs = [], [], []
    q = torch.from_numpy(xq).cuda()
    shards = []
    for r in range(G):
        s, e = pkg.shard_bounds(N, G, r)
        sh = pkg.FlatIndex(Dm, pkg.METRIC_L2, "bf16")
        sh.set_id_offset(s)
        sh.add(xb[s:e])
        sh.set_labels(labels[s:e])
        kk, gg, ll, qn = sh.search_shard(q, k)
        keys.append(kk), gids.append(gg), labs.append(ll)
        shards.append(sh)
    D, I, L = shards[0].merge_shards(torch.stack(keys, 1), torch.stack(gids, 1), torch.stack(labs, 1), qn)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(I.cpu().numpy(), If)
    np.testing.assert_array_equal(D.cpu().numpy(), Df)
    np.testing.assert_array_equal(L.cpu().numpy(), Lf)
    assert If[0, 0] == 3 and If[0, 1] == 5000


def _multi_devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 4))) if n >= 2 else [0, 0, 0]       # one GPU: three shards on the same device


@pytest.mark.parametrize("store", ["bf16", "f32"])
def test_multi_gpu_single_process_equals_one_gpu(pkg, store, tmp_path):
    """MultiGpuFlatIndex (one process drives every shard; water-filled incremental adds; peer-memory merge kernel) must
    return exactly what one FlatIndex returns: ids, distances, labels, reconstruct, and the saved faiss file."""
    import torch
    N, Dm, Q, k = 30011, 96, 70, 15
    rng = np.random.default_rng(3)
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)     # lattice: exact in every store dtype
    xb[20000] = xb[7]
    xq = rng.integers(-2, 3, size=(Q, Dm)).astype(np.float32)
    xq[0] = xb[7]                                                 # tie across shards -> lowest global id first
    labels = (np.arange(N) % 3 == 0).astype(np.float32)
    one = pkg.FlatIndex(Dm, pkg.METRIC_L2, store)
    multi = pkg.MultiGpuFlatIndex(Dm, pkg.METRIC_L2, store, devices=_multi_devices())
    for a, b in ((0, 9000), (9000, 9100), (9100, 22000), (22000, N)):      # big, tiny, big, big adds
        one.add(xb[a:b])
        multi.add(xb[a:b])
    one.set_labels(labels)
    multi.set_labels(labels)
    assert multi.ntotal == N and sum(multi.shard_sizes) == N
    assert max(multi.shard_sizes) - min(multi.shard_sizes) <= 4096 + 100   # water-filling keeps the shards balanced
    D1, I1, L1 = one.search(xq, k, return_labels=True)
    Dm_, Im_, Lm_ = multi.search(xq, k, return_labels=True)
    np.testing.assert_array_equal(Im_, I1)
    np.testing.assert_array_equal(Dm_, D1)
    np.testing.assert_array_equal(Lm_, L1)
    assert I1[0, 0] == 7 and I1[0, 1] == 20000
    Dd, Id = multi.search(torch.from_numpy(xq).cuda(), k)                  # device tensors in -> device tensors out
    torch.cuda.synchronize()
    np.testing.assert_array_equal(Id.cpu().numpy(), I1)
    np.testing.assert_array_equal(Dd.cpu().numpy(), D1)
    ids = np.array([[0, 8999, 9000], [9050, 29999, -1]], dtype=np.int64)
    np.testing.assert_array_equal(multi.reconstruct_batch(ids), one.reconstruct_batch(ids))
    np.testing.assert_array_equal(multi.reconstruct(21999), one.reconstruct(21999))
    with pytest.raises(RuntimeError):
        multi.reconstruct(N)
    # persistence: same bytes as the one-GPU writer; loads back sharded and on one GPU
    p1, pm = str(tmp_path / "one.bin"), str(tmp_path / "multi.bin")
    one.save(p1)
    multi.save(pm)
    assert open(p1, "rb").read() == open(pm, "rb").read()
    back = pkg.MultiGpuFlatIndex.load(p1, store, devices=_multi_devices())
    back.set_labels(labels)
    Db, Ib, Lb = back.search(xq, k, return_labels=True)
    np.testing.assert_array_equal(Ib, I1)
    np.testing.assert_array_equal(Db, D1)
    np.testing.assert_array_equal(Lb, L1)


def test_vector_database_on_several_gpus(pkg, tmp_path):
    """config.db_devices: the unmodified VectorDatabase / retrieve_similar_vectors surface on a row-sharded index."""
    import torch
    g = _load(os.path.join(GOLDEN, "search_gauss_l2.npz"))
    base = pkg.VectorDatabase(Cfg(tmp_path / "a", "L2", top_k=5))
    many = pkg.VectorDatabase(Cfg(tmp_path / "b", "L2", top_k=5, db_devices=_multi_devices()))
    _fill(base, g)
    _fill(many, g)
    assert type(many.index).__name__ == "MultiGpuFlatIndex" and many.index.ntotal == base.index.ntotal
    D0, I0 = base.search_batch(g["xq"], k=15)
    D1, I1 = many.search_batch(g["xq"], k=15)
    np.testing.assert_array_equal(I1, I0)
    np.testing.assert_array_equal(D1, D0)
    q = torch.from_numpy(g["xq"][:16]).cuda()
    qp = [base.vector_paths[int(i)] for i in I0[:16, 0]]
    v0, l0, p0, d0 = pkg.retrieve_similar_vectors(base, q, 5, query_paths=qp, return_info=True, return_distances=True)
    v1, l1, p1, d1 = pkg.retrieve_similar_vectors(many, q, 5, query_paths=qp, return_info=True, return_distances=True)
    assert p0 == p1
    assert torch.equal(v0, v1) and torch.equal(l0, l1) and torch.equal(d0.nan_to_num(-1), d1.nan_to_num(-1))
    many.save()
    again = pkg.VectorDatabase(Cfg(tmp_path / "b", "L2", top_k=5, db_devices=_multi_devices()))
    again.load()
    D2, I2 = again.search_batch(g["xq"], k=15)
    np.testing.assert_array_equal(I2, I0)


def test_random_shapes_lattice_fuzz(pkg, oracle):
    """Seeded fuzz over shapes / metrics / stores / k / batch sizes on lattice data (exact arithmetic in every store
    dtype): whatever scorer AUTO picks (stream, tensor cores incl. query-stationary and split-precision forms, exact
    CUDA cores) must return the oracle's ids and distances bit-for-bit, lowest id on ties.  Covers ragged N, D not a
    multiple of 8 / 64, nq around the 128-query tile boundary, k = 1, k > 32, k > N clamped by the caller."""
    rng = np.random.default_rng(20261018)
    for trial in range(36):
        N = int(rng.choice([257, 1000, 4097, 9001, 20011]))
        Dm = int(rng.choice([8, 20, 64, 100, 192, 256, 260, 520]))
        nq = int(rng.choice([1, 3, 4, 5, 64, 127, 129, 257]))
        k = int(rng.choice([1, 5, 10, 16, 17, 32, 33, 100]))
        store = str(rng.choice(["f32", "bf16", "f16"]))
        metric_s = str(rng.choice(["L2", "IP"]))
        if store == "f32" and Dm % 4 != 0 and nq <= 4:
            Dm += 4 - Dm % 4
        k = min(k, N)
        metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
        xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
        xq = rng.integers(-2, 3, size=(nq, Dm)).astype(np.float32)
        xb[N // 2] = xb[1]
        xq[0] = xb[1]
        idx = pkg.FlatIndex(Dm, metric, store)
        cut = int(rng.integers(1, N))
        idx.add(xb[:cut])
        idx.add(xb[cut:])
        D, I = idx.search(xq, k)
        ref = oracle.FlatIndexOracle(Dm, metric)
        ref.add(xb)
        Dr, Ir = ref.search(xq, k, direct=False)
        tag = f"trial {trial}: N={N} D={Dm} nq={nq} k={k} {store} {metric_s} scorer={idx.last_kernel_ms()[1]}"
        np.testing.assert_array_equal(I, Ir, err_msg=tag)
        np.testing.assert_array_equal(D, Dr, err_msg=tag)
        idx.close()


@pytest.mark.parametrize("store,algo", [("bf16", "tc"), ("f32", "tc"), ("f32", "simt"), ("bf16", "stream")])
def test_clustered_distribution_with_duplicates_and_vote(pkg, oracle, store, algo):
    """SURVEY 8(d) Dist-B: rows = centroid + 0.3 N(0,1) around 512 centroids, label = centroid id mod 2, queries =
    perturbed database rows with exact duplicates of database rows mixed in (distance 0 / self-match, the case the
    caller's exclude-self depth K + 10 exists for).  Ids within the stated tolerance, labels = labels[ids], and the kNN
    label vote equals the sum of the first K neighbour labels."""
    rng = np.random.default_rng(77)
    N, Dm, Q, k, K = 40000, 128, 4 if algo == "stream" else 300, 15, 5
    cent = rng.standard_normal((512, Dm)).astype(np.float32)
    cid = rng.integers(0, 512, size=N)
    xb = (cent[cid] + 0.3 * rng.standard_normal((N, Dm))).astype(np.float32)
    labels = (cid % 2).astype(np.float32)
    src = rng.integers(0, N, size=Q)
    xq = (xb[src] + 0.1 * rng.standard_normal((Q, Dm))).astype(np.float32)
    dup = np.arange(0, Q, 3)
    xq[dup] = xb[src[dup]]                                     # exact duplicates of database rows
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, store)
    idx.add(xb)
    idx.set_labels(labels)
    D, I, L = idx.search(xq, k, algo=algo, return_labels=True)
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2, store=store)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k + 8, direct=False)
    scale = float((xq * xq).sum(1).max() + (xb * xb).sum(1).max())
    floor = (2e-6 if (store == "f32") else 1e-4) * scale
    st = oracle.compare_topk(D, I, Dr, Ir, lambda ids: ref.exact_scores(xq, ids), pkg.METRIC_L2, tol=TOL_F32 if store == "f32" else TOL_BF16, abs_floor=floor)
    assert st["recall"] >= 0.999, st
    np.testing.assert_array_equal(I[dup, 0], src[dup])         # the duplicate finds its own row first ...
    assert float(D[dup, 0].max()) <= floor                     # ... at distance ~0
    np.testing.assert_array_equal(L, labels[I])
    vote = idx.label_vote(L, K)
    np.testing.assert_allclose(vote, labels[I][:, :K].sum(1))
    # same-cluster neighbours dominate: the vote agrees with the query's own cluster label for almost every query
    agree = np.mean((vote >= (K + 1) // 2) == (labels[src] > 0.5))
    assert agree > 0.95, agree


def test_multi_gpu_tiny_and_uneven_shards(pkg, oracle):
    """MultiGpuFlatIndex corner cases: fewer rows than shards (empty shards), shards holding fewer than k rows (their
    lists end in -1 slots), k == ntotal, labels set before further adds, and a stream-scorer sized batch (nq <= 4)."""
    rng = np.random.default_rng(9)
    Dm = 24
    xb = rng.integers(-2, 3, size=(5000, Dm)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(6, Dm)).astype(np.float32)
    multi = pkg.MultiGpuFlatIndex(Dm, pkg.METRIC_L2, "bf16", devices=_multi_devices())
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    multi.add(xb[:2])                                   # 2 rows, >= 3 shards: at least one shard stays empty
    ref.add(xb[:2])
    D, I = multi.search(xq, 2)
    Dr, Ir = ref.search(xq, 2, direct=False)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    for a, b in ((2, 9), (9, 40), (40, 41)):            # tiny adds: each goes whole to the least-loaded shard
        multi.add(xb[a:b])
        ref.add(xb[a:b])
    assert multi.ntotal == 41 and min(multi.shard_sizes) < 15
    D, I = multi.search(xq, 41)                         # k == ntotal > rows of any single shard
    Dr, Ir = ref.search(xq, 41, direct=False)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    labels = (np.arange(41) % 2).astype(np.float32)
    multi.set_labels(labels)
    D, I, L = multi.search(xq[:3], 7, return_labels=True)          # nq <= 4: streaming scorer on every shard
    np.testing.assert_array_equal(I, Ir[:3, :7])
    np.testing.assert_array_equal(L, labels[I])
    multi.add(xb[41:])                                  # big add: water-filled over all shards
    ref.add(xb[41:])
    D, I = multi.search(xq, 15)
    Dr, Ir = ref.search(xq, 15, direct=False)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    assert max(multi.shard_sizes) - min(multi.shard_sizes) <= 41


def test_reconstruct_after_search_is_served_from_one_gather(pkg):
    """index.reconstruct(i) for ids of the last host-path search (the reference caller's access pattern,
    pipeline.py:491-509) must return exactly the stored rows -- whether they come from the batched cache or from a
    single-row device read -- also for ids that were NOT in the search result, after further adds, and it must hand
    out independent copies."""
    rng = np.random.default_rng(4)
    xb = rng.standard_normal((3000, 40)).astype(np.float32)
    xq = rng.standard_normal((9, 40)).astype(np.float32)
    for store in ("f32", "bf16"):
        idx = pkg.FlatIndex(40, pkg.METRIC_L2, store)
        idx.add(xb)
        truth = idx.reconstruct_batch(np.arange(3000))
        D, I = idx.search(xq, 15)
        launches = idx.launch_count
        got = np.stack([idx.reconstruct(int(i)) for i in I.ravel()])
        np.testing.assert_array_equal(got, truth[I.ravel()])
        assert idx.launch_count - launches <= 1                     # ONE gather kernel for all 135 calls
        a = idx.reconstruct(int(I[0, 0]))
        a[:] = 0
        np.testing.assert_array_equal(idx.reconstruct(int(I[0, 0])), truth[I[0, 0]])
        miss = int(np.setdiff1d(np.arange(3000), I.ravel())[0])
        np.testing.assert_array_equal(idx.reconstruct(miss), truth[miss])
        idx.add(xb[:10] + 1.0)
        np.testing.assert_array_equal(idx.reconstruct(3005), idx.reconstruct_batch(np.array([3005]))[0])
        with pytest.raises(RuntimeError):
            idx.reconstruct(99999)


def test_concurrent_searches_from_threads(pkg, oracle):
    """Several Python threads searching the same index and their own indexes at once (a Flask dev server can call
    predict concurrently, app.py:351): every handle serialises internally, ctypes releases the GIL inside the C call,
    and every result must equal the single-threaded one."""
    import threading
    rng = np.random.default_rng(8)
    xb = rng.integers(-2, 3, size=(20000, 64)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(64, 64)).astype(np.float32)
    shared = pkg.FlatIndex(64, pkg.METRIC_L2, "bf16")
    shared.add(xb)
    Dref, Iref = shared.search(xq, 10)
    D1ref, I1ref = shared.search(xq[:1], 10)
    errors = []

    def worker(t):
        try:
            own = pkg.FlatIndex(64, pkg.METRIC_L2, "bf16")
            own.add(xb)
            for it in range(15):
                for index in (shared, own):
                    D, I = index.search(xq, 10)
                    assert np.array_equal(I, Iref) and np.array_equal(D, Dref)
                    D1, I1 = index.search(xq[:1], 10)                   # streaming scorer (ticket / control block)
                    assert np.array_equal(I1, I1ref) and np.array_equal(D1, D1ref)
            own.close()
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("store", ["bf16", "f32"])
def test_more_queries_than_one_internal_batch(pkg, oracle, store):
    """nq > 65 536 is processed in internal batches (scratch reuse between batches, host and device paths)."""
    import torch
    rng = np.random.default_rng(12)
    N, Dm, nq, k = 3000, 16, 70001, 5
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(nq, Dm)).astype(np.float32)
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, store)
    idx.add(xb)
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    D, I = idx.search(xq, k)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    Dd, Id = idx.search(torch.from_numpy(xq).cuda(), k)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(Id.cpu().numpy(), Ir)
    np.testing.assert_array_equal(Dd.cpu().numpy(), Dr)
