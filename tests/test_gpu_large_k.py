"""GPU parity for 128 < k <= 2048 (index.search(q, k) -- vector_database.py:181 -- accepts any k; faiss-gpu up to 2048).

The path: dense fp32 keys of (query block x row chunk) written to HBM -- by the tensor-core scorer's SelectDump epilogue
for 16-bit stores, by the exact CUDA-core scorer's DUMP form for fp32 stores -> exact radix select per query and chunk
(csrc/select_large.cuh) -> merge of the per-chunk lists.  Integer-lattice inputs (exact
arithmetic, massive ties) must match the oracle BIT-EXACTLY, ties by the lowest id; Gaussian inputs within the
north-star tolerances (1e-5 relative fp32 store, 1e-3 16-bit stores).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_F32, TOL_BF16 = 1e-5, 1e-3


def _gauss(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def _lattice(N, D, Q, seed):
    rng = np.random.default_rng(seed)
    xb = rng.integers(-2, 3, size=(N, D)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(Q, D)).astype(np.float32)
    if N > 700:
        xb[N // 2:N // 2 + 300] = xb[10]                        # 300 identical rows: ties far beyond the boundary
        xq[0] = xb[10]
    return xb, xq


@pytest.mark.parametrize("store,scorer", [("f32", "simt"), ("f32", "split"), ("bf16", "tc"), ("bf16", "simt"), ("f16", "tc")])
@pytest.mark.parametrize("metric_s", ["L2", "IP"])
@pytest.mark.parametrize("k,rows", [(129, None), (200, 1024), (777, None), (2048, 2500)])
def test_lattice_bit_exact_k_above_128(pkg, oracle, metric_s, k, rows, store, scorer):
    """rows = forced chunk length (option "largek_rows"): several per-chunk lists, ties straddling chunk borders.
    scorer = where the dense keys come from: the tensor cores (16-bit stores; fp32 stores: split precision + certified
    exact re-rank) or the exact CUDA-core kernel."""
    xb, xq = _lattice(6001, 64, 70, 11)
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(64, metric, store)
    # fp32 stores: "simt" = exact CUDA-core dense keys, "split" = split-precision tensor-core keys + exact re-rank +
    # certificate (the 300 identical rows outnumber the candidate slack, so those queries also walk the exact fallback)
    idx.set_option("largek_scorer", {"simt": 1, "tc": 2, "split": 0}[scorer])
    idx.set_option("largek_split", 1 if scorer == "split" else 0)
    if rows:
        idx.set_option("largek_rows", rows)
    idx.add(xb[:4000])
    idx.add(xb[4000:])
    D, I = idx.search(xq, k)
    ref = oracle.FlatIndexOracle(64, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (70, k)
    assert idx.last_kernel_ms()[1] in {"split": ("tc", "simt")}.get(scorer, (scorer,))
    if scorer == "split":
        assert idx.last_uncertified >= 1            # the tie block cannot be certified: exact fallback exercised


def test_all_rows_identical(pkg):
    """Every key equal: the select has to walk all eight radix digits down to the row id."""
    xb = np.ones((5000, 16), np.float32)
    idx = pkg.FlatIndex(16, pkg.METRIC_L2, "f32")
    idx.add(xb)
    D, I = idx.search(np.ones((3, 16), np.float32), 1500)
    np.testing.assert_array_equal(I, np.tile(np.arange(1500, dtype=np.int64), (3, 1)))
    np.testing.assert_array_equal(D, np.zeros((3, 1500), np.float32))


@pytest.mark.parametrize("case", [
    # N      D    Q    k     metric cos    store   rows
    (20000,  128, 300, 1000, "L2", False, "f32",  None),      # two query blocks of 256
    (30011,  96,  64,  500,  "IP", True,  "bf16", 8192),      # 4 chunks, ragged last chunk
    (5003,   101, 33,  256,  "IP", False, "f32",  None),      # D % 4 != 0
    (9000,   768, 1,   150,  "L2", False, "bf16", None),      # batch-1 (beyond the streaming scorer's k)
    (4100,   40,  5,   2048, "L2", False, "f16",  None),
], ids=["l2_f32_k1000", "cos_bf16_k500_chunks", "ip_f32_oddD_k256", "l2_bf16_q1_k150", "l2_f16_k2048"])
def test_gaussian_vs_oracle_k_above_128(pkg, oracle, case):
    N, Dm, Q, k, metric_s, cos, store, rows = case
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    xb, xq = _gauss(N, Dm, 1234), _gauss(Q, Dm, 5678)
    xq[0] = xb[5]
    idx = pkg.FlatIndex(Dm, metric, store)
    if rows:
        idx.set_option("largek_rows", rows)
    idx.add(xb, normalize=cos)
    D, I = idx.search(xq, k, normalize=cos)
    ref = oracle.FlatIndexOracle(Dm, metric, store=store)
    ref.add(oracle.maybe_normalize(xb, cos))
    qn = oracle.maybe_normalize(xq, cos)
    Dr, Ir = ref.search(qn, min(k + 8, N), direct=False)
    tol = TOL_F32 if store == "f32" else TOL_BF16
    if metric == pkg.METRIC_L2:
        scale = float((qn * qn).sum(1).max() + (ref._base() ** 2).sum(1).max())
        floor = (2e-6 if store == "f32" else 1e-4) * scale       # 16-bit stores: keys from the tensor cores
    else:
        floor = 1e-6
    st = oracle.compare_topk(D, I, Dr, Ir, lambda ids: ref.exact_scores(qn, ids), metric, tol=tol, abs_floor=floor)
    assert st["recall"] >= 0.999, st
    # fp32 stores: split-precision tensor-core keys + certified exact re-rank (exact CUDA-core keys when D % 4 != 0)
    assert idx.last_kernel_ms()[1] == ("simt" if (store == "f32" and Dm % 4) else "tc")
    # sorted best-first, no duplicate ids
    assert np.all(np.diff(D, axis=1) >= 0) if metric == pkg.METRIC_L2 else np.all(np.diff(D, axis=1) <= 0)
    assert all(len(set(r.tolist())) == k for r in I)


@pytest.mark.parametrize("sample", [True, False])
@pytest.mark.parametrize("kind", ["gauss", "lattice", "on_stride"])
def test_long_rows_sampled_pivot_select(pkg, oracle, kind, sample):
    """Chunks of >= 32768 rows take the select's fast path (pivot from a strided sample + one collect pass); it must be
    exact on Gaussian keys, on lattice keys (ties ordered by id), and when the sample misleads (the best rows sit
    exactly on the sample stride -> too few rows reach the pivot -> exact path).  Same answers with the fast path off."""
    N, Dm, Q, k = 163_840, 32, 12, 700
    rng = np.random.default_rng(8)
    if kind == "gauss":
        xb, xq = _gauss(N, Dm, 1), _gauss(Q, Dm, 2)
    else:
        xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
        xq = rng.integers(-2, 3, size=(Q, Dm)).astype(np.float32)
        if kind == "on_stride":
            xb[::10] = xq[0]                                      # stride of the sample = N // 16384 = 10
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "bf16" if kind != "gauss" else "f32")
    idx.set_option("largek_sample", 1 if sample else 0)
    idx.add(xb)
    D, I = idx.search(xq, k)
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    ref.add(xb)
    if kind == "gauss":
        Dr, Ir = ref.search(xq, k + 8, direct=False)
        scale = float((xq * xq).sum(1).max() + (xb * xb).sum(1).max())
        st = oracle.compare_topk(D, I, Dr, Ir, lambda ids: ref.exact_scores(xq, ids), pkg.METRIC_L2, tol=TOL_F32,
                                 abs_floor=2e-6 * scale)
        assert st["recall"] == 1.0, st
    else:
        Dr, Ir = ref.search(xq, k, direct=False)
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D, Dr)


def test_k_beyond_ntotal_and_limits(pkg, tmp_path):
    """k > ntotal at the index level: faiss fills id -1 / +inf; the wrapper clamps k to ntotal (vector_database.py:169);
    k > 2048 is refused loudly (faiss-gpu's own limit)."""
    from conftest import Cfg
    xb = _gauss(200, 32, 3)
    idx = pkg.FlatIndex(32, pkg.METRIC_L2, "f32")
    idx.add(xb)
    D, I = idx.search(xb[:4], 300)
    assert np.all(I[:, 200:] == -1) and np.all(np.isinf(D[:, 200:]))
    assert all(sorted(r[:200].tolist()) == list(range(200)) for r in I)
    np.testing.assert_array_equal(I[:, 0], np.arange(4))
    with pytest.raises(RuntimeError):
        idx.search(xb[:4], 2049)
    vdb = pkg.VectorDatabase(Cfg(tmp_path / "lk", "L2"))
    vdb.add_vectors(xb, [f"/d/u{i}.wav" for i in range(200)], [i % 2 for i in range(200)], {})
    Dw, Iw = vdb.search_batch(xb[:4], k=1000)
    assert Dw.shape == (4, 200) and Iw.shape == (4, 200)
    np.testing.assert_array_equal(Iw, I[:, :200])


def test_sharded_and_labels_k_above_128(pkg, oracle):
    """Row shards (4 emulated on one GPU) + merge kernel == unsharded search at k = 400, labels gathered on device."""
    import torch
    N, Dm, Q, k = 10007, 64, 37, 400
    xb, xq = _lattice(N, Dm, Q, 5)
    labels = (np.arange(N) % 2).astype(np.float32)
    full = pkg.FlatIndex(Dm, pkg.METRIC_L2, "bf16")
    full.add(xb)
    full.set_labels(labels)
    Df, If, Lf = full.search(xq, k, return_labels=True)
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    np.testing.assert_array_equal(If, Ir)
    np.testing.assert_array_equal(Df, Dr)
    np.testing.assert_array_equal(Lf, labels[If])
    G = 4
    q = torch.from_numpy(xq).cuda()
    keys, gids, labs, shards = [], [], [], []
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


def test_multi_gpu_index_k_above_128(pkg):
    """MultiGpuFlatIndex (peer-memory merge kernel) at k = 300 == one FlatIndex."""
    import torch
    n = torch.cuda.device_count()
    devices = list(range(min(n, 4))) if n >= 2 else [0, 0, 0]
    xb, xq = _lattice(9001, 64, 20, 17)
    one = pkg.FlatIndex(64, pkg.METRIC_IP, "bf16")
    one.add(xb)
    D1, I1 = one.search(xq, 300)
    multi = pkg.MultiGpuFlatIndex(64, pkg.METRIC_IP, "bf16", devices=devices)
    multi.add(xb[:5000])
    multi.add(xb[5000:])
    Dm_, Im_ = multi.search(xq, 300)
    np.testing.assert_array_equal(Im_, I1)
    np.testing.assert_array_equal(Dm_, D1)
