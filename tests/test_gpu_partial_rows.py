"""Rows of SURVEY section 8 the round-1 review marked partial: fp32 master copy next to a 16-bit store (f2), the on-disk
format beyond the happy path (f3), ingest from device tensors through VectorDatabase (a14 / f4), resource release (a11),
and the ADVICE.md findings (global row bound of the first-K filter on row-sharded indexes, odd nq*k in the peer-exchange
buffer, atomic multi-shard add, label mismatch)."""
import os
import struct

import numpy as np
import pytest

from conftest import Cfg

pytestmark = pytest.mark.gpu


def _gauss(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ f2
@pytest.mark.parametrize("store", ["bf16", "f16"])
def test_keep_f32_master_exact_reconstruct_with_16bit_scoring(pkg, oracle, store, tmp_path):
    """db_keep_f32_master: neighbours are scored on the 16-bit rows (same ids / distances as a plain 16-bit store), but
    index.reconstruct (pipeline.py:503) returns the fp32 rows exactly -- what identical logits need."""
    xb, xq = _gauss(3000, 96, 1), _gauss(33, 96, 2)
    plain = pkg.FlatIndex(96, pkg.METRIC_L2, store)
    keep = pkg.FlatIndex(96, pkg.METRIC_L2, store, keep_f32_master=True)
    plain.add(xb)
    keep.add(xb)
    Dp, Ip = plain.search(xq, 15)
    Dk, Ik = keep.search(xq, 15)
    np.testing.assert_array_equal(Ip, Ik)
    np.testing.assert_array_equal(Dp, Dk)
    ids = Ik[:, :5].reshape(-1)
    np.testing.assert_array_equal(keep.reconstruct_batch(ids), xb[ids])                  # exact fp32
    np.testing.assert_array_equal(keep.reconstruct(int(ids[3])), xb[ids[3]])
    rounded = oracle.round_bf16(xb) if store == "bf16" else oracle.round_fp16(xb)
    np.testing.assert_array_equal(plain.reconstruct_batch(ids), rounded[ids])            # faiss useFloat16 behaviour
    assert keep.mem_info()["index_bytes"] > plain.mem_info()["index_bytes"]
    # through the wrapper: config.db_keep_f32_master, incl. save -> load (the file holds the fp32 rows)
    cfg = Cfg(tmp_path / "m", "L2", db_dtype=store, db_keep_f32_master=True)
    vdb = pkg.VectorDatabase(cfg)
    vdb.add_vectors(xb, [f"p{i}" for i in range(len(xb))], [0] * len(xb), {})
    np.testing.assert_array_equal(vdb.index.reconstruct(17), xb[17])
    vdb.save()
    v2 = pkg.VectorDatabase(cfg)
    v2.load()
    np.testing.assert_array_equal(v2.index.reconstruct(17), xb[17])
    np.testing.assert_array_equal(v2.search_batch(xq, k=15)[1], Ik)


# ------------------------------------------------------------------------------------------------ f3
def _faiss_header(cc, d, n, metric, count=None, metric_arg=None):
    h = cc + struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric)
    if metric_arg is not None:
        h += struct.pack("<f", metric_arg)
    return h + struct.pack("<Q", n * d if count is None else count)


def test_on_disk_format_variants(pkg, tmp_path):
    xb = _gauss(41, 12, 3)
    # IxF2 written by us, byte for byte the faiss layout
    idx = pkg.FlatIndex(12, pkg.METRIC_L2, "f32")
    idx.add(xb)
    p = str(tmp_path / "l2.bin")
    idx.save(p)
    raw = open(p, "rb").read()
    assert raw == _faiss_header(b"IxF2", 12, 41, 1) + xb.tobytes()
    # a foreign writer's files: IxF2 / IxFI / the legacy generic IxFl fourcc
    for cc, metric, want in ((b"IxF2", 1, pkg.METRIC_L2), (b"IxFI", 0, pkg.METRIC_IP), (b"IxFl", 1, pkg.METRIC_L2)):
        q = str(tmp_path / f"foreign_{cc.decode()}.bin")
        open(q, "wb").write(_faiss_header(cc, 12, 41, metric) + xb.tobytes())
        back = pkg.FlatIndex.load(q, "f32")
        assert (back.ntotal, back.d, back.metric) == (41, 12, want)
        np.testing.assert_array_equal(back.reconstruct_batch(np.arange(41)), xb)
    # metric types > 1 carry a metric_arg float in faiss files: parsed, then refused (only L2 / IP are flat-searchable here)
    q = str(tmp_path / "lp.bin")
    open(q, "wb").write(_faiss_header(b"IxFl", 12, 41, 4, metric_arg=3.0) + xb.tobytes())
    with pytest.raises(RuntimeError, match="metric"):
        pkg.FlatIndex.load(q, "f32")
    # truncated body / truncated header / wrong count / not an index at all
    for name, blob in (("trunc_body", raw[:-100]), ("trunc_head", raw[:20]),
                       ("bad_count", _faiss_header(b"IxF2", 12, 41, 1, count=7) + xb.tobytes()),
                       ("foreign", b"IwFl" + raw[4:]), ("garbage", os.urandom(64)), ("empty", b"")):
        q = str(tmp_path / f"{name}.bin")
        open(q, "wb").write(blob)
        with pytest.raises(RuntimeError):
            pkg.FlatIndex.load(q, "f32")
        with pytest.raises(RuntimeError):
            pkg.MultiGpuFlatIndex.load(q, "f32", devices=[0, 0])
    # the wrapper's load() never raises (vector_database.py:241-242): a corrupt file leaves index None
    cfg = Cfg(tmp_path / "w", "L2")
    vdb = pkg.VectorDatabase(cfg)
    vdb.add_vectors(xb, [f"p{i}" for i in range(41)], [0] * 41, {})
    vdb.save()
    open(vdb.db_path, "wb").write(raw[:50])
    v2 = pkg.VectorDatabase(cfg)
    v2.load()
    assert v2.index is None
    # empty index round trip
    e = pkg.FlatIndex(5, pkg.METRIC_IP, "f32")
    q = str(tmp_path / "empty.idx")
    e.save(q)
    assert pkg.FlatIndex.load(q, "bf16").ntotal == 0


# ------------------------------------------------------------------------------------------------ a14 / f4
@pytest.mark.parametrize("itype", ["L2", "IP"])
def test_vector_database_accepts_cuda_tensors(pkg, itype, tmp_path):
    """build_vector_database (pipeline.py:416-447) can hand the encoder's CUDA tensors straight to add_vectors: same
    database, bit for bit, as the numpy route (`.cpu().numpy()` + np.vstack, pipeline.py:430,444)."""
    import torch
    xb, xq = _gauss(2500, 64, 5), _gauss(20, 64, 6)
    paths = [f"/d/f{i}.wav" for i in range(len(xb))]
    labels = [i % 2 for i in range(len(xb))]
    meta = {"speaker_id": [f"s{i % 7}" for i in range(len(xb))], "split": "train"}
    a = pkg.VectorDatabase(Cfg(tmp_path / "a", itype, vector_add_batch_size=700))
    b = pkg.VectorDatabase(Cfg(tmp_path / "b", itype, vector_add_batch_size=700))
    a.add_vectors(xb, paths, labels, meta)
    b.add_vectors(torch.from_numpy(xb).cuda(), paths, labels, meta)
    assert b.index.ntotal == len(xb) and b.vector_paths == a.vector_paths and b.vector_labels == a.vector_labels
    assert b.vector_metadata == a.vector_metadata
    np.testing.assert_array_equal(b.index.reconstruct_batch(np.arange(len(xb))), a.index.reconstruct_batch(np.arange(len(xb))))
    Da, Ia = a.search_batch(xq, k=15)
    Db, Ib = b.search_batch(torch.from_numpy(xq).cuda(), k=15)
    np.testing.assert_array_equal(Ib.cpu().numpy(), Ia)
    np.testing.assert_array_equal(Db.cpu().numpy(), Da)
    # fp16 / fp64 tensors and non-contiguous views are converted like the numpy route's astype(float32)
    c = pkg.VectorDatabase(Cfg(tmp_path / "c", itype))
    c.add_vectors(torch.from_numpy(xb).cuda().double()[::2], paths[::2], labels[::2], {})
    np.testing.assert_array_equal(c.index.reconstruct(3), a.index.reconstruct(6))


# ------------------------------------------------------------------------------------------------ a11
def test_cleanup_releases_device_memory(pkg, tmp_path):
    xb = _gauss(200_000, 128, 9)
    cfg = Cfg(tmp_path / "c", "L2")
    vdb = pkg.VectorDatabase(cfg)
    free0 = vdb.get_gpu_memory_usage()["total"] - vdb.get_gpu_memory_usage()["used"]
    vdb.add_vectors(xb, [""] * len(xb), [0] * len(xb), {})
    vdb.search_batch(xb[:300], k=10)
    idx = vdb.index
    info = idx.mem_info()
    assert info["index_bytes"] >= xb.nbytes and info["scratch_bytes"] > 0
    vdb.save()
    vdb.cleanup_gpu_resources()                                    # vector_database.py:259-268
    assert vdb.index is None and vdb.gpu_index is None
    after = idx.mem_info()
    assert after["index_bytes"] == 0 and after["scratch_bytes"] == 0
    assert after["free"] >= info["free"] + info["index_bytes"] // 2          # the device really got it back
    with pytest.raises(ValueError, match="Vector database is empty"):        # reference text (vector_database.py:161)
        vdb.search_batch(xb[:3], k=3)
    vdb.cleanup_gpu_resources()                                    # idempotent
    vdb.load()                                                     # and the database comes back from disk
    assert vdb.index.ntotal == len(xb)
    assert vdb.search_batch(xb[:3], k=1)[1].ravel().tolist() == [0, 1, 2]
    # release_index=False keeps the index searchable and only drops the grow-only scratch (faiss's StandardGpuResources)
    vdb.cleanup_gpu_resources(release_index=False)
    assert vdb.index.mem_info()["scratch_bytes"] == 0 and vdb.index.ntotal == len(xb)
    assert vdb.search_batch(xb[:3], k=1)[1].ravel().tolist() == [0, 1, 2]
    del vdb                                                        # __del__ -> cleanup (vector_database.py:270-273)
    assert free0 > 0


# ------------------------------------------------------------------------------------------------ ADVICE.md
def test_retrieve_on_row_sharded_index_keeps_neighbours_of_every_shard(pkg, tmp_path):
    """ADVICE high: the first-K filter used ONE shard's row count as the id bound, dropping ~(G-1)/G of the neighbours of
    a row-sharded database (N >= 4096 * G so that every shard really holds rows)."""
    import torch
    G = 3
    N, D, B, K = 4096 * G + 500, 48, 64, 5
    xb = _gauss(N, D, 11)
    paths = [f"/d/spk{i % 11}/utt_{i:06d}.wav" for i in range(N)]
    labels = [(i * 7) % 2 for i in range(N)]
    one = pkg.VectorDatabase(Cfg(tmp_path / "one", "L2", top_k=K, vector_add_batch_size=N))
    multi = pkg.VectorDatabase(Cfg(tmp_path / "multi", "L2", top_k=K, db_devices=[0] * G, vector_add_batch_size=N))
    one.add_vectors(xb, paths, labels, {})
    multi.add_vectors(xb, paths, labels, {})
    assert min(multi.index.shard_sizes) > 4000
    rng = np.random.default_rng(12)
    src = rng.integers(0, N, size=B)
    q = torch.from_numpy(xb[src] + 0.01 * _gauss(B, D, 13)).cuda()
    qpaths = [paths[i] for i in src]
    v1, l1, p1, d1 = pkg.retrieve_similar_vectors(one, q, K, D, query_paths=qpaths, return_info=True, return_distances=True)
    v2, l2, p2, d2 = pkg.retrieve_similar_vectors(multi, q, K, D, query_paths=qpaths, return_info=True, return_distances=True)
    assert p1 == p2 and all(all(s != "" for s in row) for row in p2)
    np.testing.assert_array_equal(v2.cpu().numpy(), v1.cpu().numpy())
    np.testing.assert_array_equal(l2.cpu().numpy(), l1.cpu().numpy())
    np.testing.assert_array_equal(d2.cpu().numpy(), d1.cpu().numpy())
    hit_shards = {int(np.searchsorted(np.cumsum(multi.index.shard_sizes), paths.index(s), side="right"))
                  for row in p2[:16] for s in row}
    assert len(hit_shards) == G                                    # neighbours really come from every shard


def test_multi_gpu_add_is_atomic(pkg):
    """ADVICE medium: a failure on a later shard must not leave rows (and global ids) behind on the earlier ones."""
    D = 32
    multi = pkg.MultiGpuFlatIndex(D, pkg.METRIC_L2, "f32", devices=[0, 0, 0])
    xb = _gauss(13000, D, 21)
    multi.add(xb[:6000])
    sizes = multi.shard_sizes
    calls = {"n": 0}
    real_add = multi.shards[2].add

    def failing_add(x, normalize=False):
        calls["n"] += 1
        raise RuntimeError("injected failure on the last shard")
    multi.shards[2].add = failing_add
    with pytest.raises(RuntimeError, match="injected"):
        multi.add(xb[6000:])
    assert calls["n"] == 1 and multi.ntotal == 6000 and multi.shard_sizes == sizes
    multi.shards[2].add = real_add
    multi.add(xb[6000:])                                           # ids continue where they should
    one = pkg.FlatIndex(D, pkg.METRIC_L2, "f32")
    one.add(xb)
    q = _gauss(50, D, 22)
    D1, I1 = one.search(q, 10)
    D2, I2 = multi.search(q, 10)
    np.testing.assert_array_equal(I2, I1)
    np.testing.assert_array_equal(D2, D1)
    np.testing.assert_array_equal(multi.reconstruct_batch(I1[:, 0]), xb[I1[:, 0]])


def test_label_mismatch_is_not_silent(pkg, tmp_path):
    """ADVICE low: labels that cannot be synced (len(vector_labels) != ntotal) must not turn into label 0."""
    xb = _gauss(500, 16, 31)
    vdb = pkg.VectorDatabase(Cfg(tmp_path / "l", "L2"))
    vdb.add_vectors(xb, [f"p{i}" for i in range(500)], [1] * 500, {})
    D, I, L = vdb.search_batch_with_labels(xb[:4], k=3)
    assert float(L.min()) == 1.0
    vdb.vector_labels = vdb.vector_labels[:100]                    # e.g. a metadata.pkl that does not match the index
    with pytest.raises(IndexError):                                 # what the reference's vector_labels[ii] would raise
        vdb.search_batch_with_labels(xb[400:404], k=3)
