"""CPU: the oracle (oracle/flat_oracle.py) against the golden vectors produced by the reference's own
vector_database.py / pipeline.py source (tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, Cfg

SEARCH = sorted(glob.glob(os.path.join(GOLDEN, "search_*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def _build(oracle, g, tmp_path, store="f32"):
    cfg = Cfg(tmp_path / "o", str(g["index_type"]), normalize_for_ip=bool(g["normalize_for_ip"]),
              vector_add_batch_size=256)
    vdb = oracle.OracleVectorDatabase(cfg, store=store)
    n = g["xb"].shape[0]
    vdb.add_vectors(g["xb"], [f"/data/spk{i % 13}/utt_{i:05d}.wav" for i in range(n)],
                    [int(v) for v in g["labels"]], {"speaker_id": [f"spk{i % 13}" for i in range(n)]})
    return vdb


def test_fixtures_present():
    assert len(SEARCH) >= 9


@pytest.mark.parametrize("path", SEARCH, ids=[os.path.basename(p)[7:-4] for p in SEARCH])
def test_oracle_matches_reference_wrapper(oracle, path, tmp_path):
    g = _load(path)
    vdb = _build(oracle, g, tmp_path)
    k = int(g["k"])
    D, I = vdb.search_batch(g["xq"], k=k)
    assert D.dtype == np.float32 and I.dtype == np.int64
    assert D.shape == g["dist"].shape and I.shape == g["idx"].shape          # k clamped to ntotal
    assert vdb.index.ntotal == int(g["ntotal"]) and bool(vdb._cosine) == bool(g["cosine"])
    assert len(vdb.vector_paths) == int(g["n_paths"]) and len(vdb.vector_metadata["speaker_id"]) == int(g["n_meta"])
    metric = oracle.METRIC_IP if str(g["index_type"]) == "IP" else oracle.METRIC_L2
    qn = vdb._maybe_normalize(g["xq"].astype(np.float32))
    name = os.path.basename(path)
    if "lattice" in name or "kat_tiny" in name:
        # integer lattice: arithmetic is exact in fp32 -> bit-exact distances and ids (ties -> lowest id)
        np.testing.assert_array_equal(I, g["idx"])
        np.testing.assert_array_equal(D, g["dist"])
    else:
        st = oracle.compare_topk(D, I, g["dist"], g["idx"], lambda ids: vdb.index.exact_scores(qn, ids), metric,
                                 tol=1e-5, abs_floor=2e-5 if metric == oracle.METRIC_L2 else 1e-6)
        assert st["recall"] == 1.0
    d1, i1 = vdb.search(g["xq"][0], k=k)
    np.testing.assert_array_equal(i1, g["idx_single"])
    dd, di = vdb.search_batch(g["xq"][:2])
    assert dd.shape == g["dist_default"].shape
    np.testing.assert_array_equal(di, g["idx_default"])
    rec = np.stack([vdb.index.reconstruct(int(i)) for i in I[0]])
    np.testing.assert_allclose(rec, g["recon_row0"], rtol=1e-6, atol=1e-7)


def test_oracle_direct_and_blas_paths_agree(oracle):
    g = _load(os.path.join(GOLDEN, "search_gauss_l2.npz"))
    idx = oracle.FlatIndexOracle(g["xb"].shape[1], oracle.METRIC_L2)
    idx.add(g["xb"])
    Da, Ia = idx.search(g["xq"], 15, direct=False)
    Db, Ib = idx.search(g["xq"], 15, direct=True)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_allclose(Da, Db, rtol=2e-5, atol=2e-4)


def test_wrapper_behaviour(oracle, tmp_path):
    with open(os.path.join(GOLDEN, "wrapper_behaviour.json")) as f:
        beh = json.load(f)
    cfg = Cfg(tmp_path / "w", "L2")
    vdb = oracle.OracleVectorDatabase(cfg)
    with pytest.raises(ValueError) as e:
        vdb.search_batch(np.zeros((1, 8), np.float32))
    assert str(e.value) == beh["empty_search_error"]
    vdb.add_vectors(np.zeros((0, 8), np.float32), [], [], {})
    assert (vdb.index is None) == beh["index_none_after_empty_add"]
    xb = np.random.default_rng(1).standard_normal((10, 8)).astype(np.float32)
    vdb.add_vectors_batch(xb, [f"p{i}" for i in range(10)], list(range(10)),
                          {"split": 7, "speaker_id": [f"s{i}" for i in range(10)]}, batch_size=4)
    assert vdb.vector_metadata["split"] == beh["meta_split"]
    assert vdb.vector_metadata["speaker_id"] == beh["meta_speaker"]
    assert vdb.vector_labels == beh["labels"]
    d0, i0 = vdb.search_batch(xb[:2], k=0)
    assert [list(d0.shape), list(i0.shape)] == beh["k0_shapes"]
    assert [str(d0.dtype), str(i0.dtype)] == beh["k0_dtypes"]
    vdb.save()
    v2 = oracle.OracleVectorDatabase(cfg)
    v2.load()
    assert v2.index.ntotal == beh["loaded_ntotal"] and hasattr(v2, "_cosine") == beh["loaded_has_cosine_attr"]
    with pytest.raises(ValueError) as e:
        oracle.OracleVectorDatabase(Cfg(tmp_path / "b", "HNSW")).create_index(8)
    assert str(e.value) == beh["bad_type_error"]


def test_cosine_after_load_quirk(oracle, tmp_path):
    g = _load(os.path.join(GOLDEN, "quirk_cosine_after_load.npz"))
    cfg = Cfg(tmp_path / "q", "IP")
    vdb = oracle.OracleVectorDatabase(cfg)
    vdb.add_vectors(g["xb"], [f"p{i}" for i in range(200)], [0] * 200, {})
    d_b, i_b = vdb.search_batch(g["xq"], k=5)
    vdb.save()
    v2 = oracle.OracleVectorDatabase(cfg)
    v2.load()
    d_a, i_a = v2.search_batch(g["xq"], k=5)
    np.testing.assert_array_equal(i_b, g["i_before"])
    np.testing.assert_array_equal(i_a, g["i_after"])
    np.testing.assert_allclose(d_b, g["d_before"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d_a, g["d_after"], rtol=1e-5, atol=1e-6)     # un-normalised queries after load()


@pytest.mark.parametrize("name", ["retrieve_l2", "retrieve_cos"])
def test_retrieve_similar_vectors_oracle(oracle, name, tmp_path):
    g = _load(os.path.join(GOLDEN, f"{name}.npz"))
    K, D = int(g["K"]), g["xb"].shape[1]
    cfg = Cfg(tmp_path / "r", str(g["index_type"]), top_k=K)
    vdb = oracle.OracleVectorDatabase(cfg)
    paths = [str(p) for p in g["paths"]]
    vdb.add_vectors(g["xb"], paths, [int(l) for l in g["labels"]], {"speaker_id": ["s"] * len(paths)})
    qpaths = [str(p) for p in g["qpaths"]]
    train_ids = {str(s) for s in g["train_ids"]}
    for tag, kw in (("excl_paths", dict(query_paths=qpaths, exclude_self=True)),
                    ("excl_train", dict(query_paths=None, exclude_self=True, training_file_ids=train_ids)),
                    ("noexcl", dict(query_paths=qpaths, exclude_self=False))):
        vec, lbl, pth, dst = oracle.retrieve_similar_vectors_oracle(vdb, g["q"], K, D, **kw)
        assert [list(r) for r in pth] == [list(map(str, r)) for r in g[f"{tag}_paths"]]
        np.testing.assert_array_equal(lbl, g[f"{tag}_lbl"])
        np.testing.assert_allclose(vec, g[f"{tag}_vec"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(dst, g[f"{tag}_dist"], rtol=1e-4, atol=2e-4, equal_nan=True)


def test_round_bf16_matches_torch(oracle):
    import torch
    x = np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 37.0
    x[:4] = [0.0, -0.0, 1.0000001, 3.3895314e38]
    ref = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oracle.round_bf16(x), ref)
