"""CPU property tests (hypothesis) of the oracle itself: it is the checker for every GPU parity claim, so its own
invariants are pinned here -- exact equality with a float64 lexsort on lattice data (ties -> lowest id), shard
invariance, k clamping, BLAS-expansion vs direct agreement."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def _lattice(rng, n, d):
    return rng.integers(-2, 3, size=(n, d)).astype(np.float32)


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 300), d=st.integers(1, 40), nq=st.integers(1, 25), k=st.integers(1, 40),
       ip=st.booleans(), seed=st.integers(0, 10_000), dup=st.booleans())
def test_oracle_equals_float64_lexsort_on_lattice(oracle, n, d, nq, k, ip, seed, dup):
    rng = np.random.default_rng(seed)
    xb, xq = _lattice(rng, n, d), _lattice(rng, nq, d)
    if dup and n > 3:
        xb[n // 2] = xb[0]
        xq[0] = xb[0]
    metric = oracle.METRIC_IP if ip else oracle.METRIC_L2
    idx = oracle.FlatIndexOracle(d, metric, block_rows=64)
    idx.add(xb[: n // 2])
    idx.add(xb[n // 2:])
    D, I = idx.search(xq, k)
    kk = min(k, n)
    x64, q64 = xb.astype(np.float64), xq.astype(np.float64)
    for r in range(nq):
        s = x64 @ q64[r] if ip else ((x64 - q64[r]) ** 2).sum(1)
        order = np.lexsort((np.arange(n), -s if ip else s))[:kk]
        assert (I[r, :kk] == order).all()
        assert (D[r, :kk] == s[order].astype(np.float32)).all()
        assert (I[r, kk:] == -1).all()


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 400), g=st.integers(2, 5), k=st.integers(1, 12), seed=st.integers(0, 10_000), ip=st.booleans())
def test_oracle_shard_invariance(oracle, pkg, n, g, k, seed, ip):
    """top-k(union) == merge of per-shard top-k under (distance, id) order -- the identity the multi-GPU path uses."""
    rng = np.random.default_rng(seed)
    d = 16
    xb, xq = _lattice(rng, n, d), _lattice(rng, 7, d)
    metric = oracle.METRIC_IP if ip else oracle.METRIC_L2
    full = oracle.FlatIndexOracle(d, metric)
    full.add(xb)
    Df, If = full.search(xq, k)
    cd, ci = [], []
    for r in range(g):
        s, e = pkg.shard_bounds(n, g, r)
        sh = oracle.FlatIndexOracle(d, metric)
        if e > s:
            sh.add(xb[s:e])
        Ds, Is = sh.search(xq, k)
        cd.append(Ds)
        ci.append(np.where(Is >= 0, Is + s, -1))
    cd, ci = np.concatenate(cd, 1), np.concatenate(ci, 1)
    kk = min(k, n)
    for q in range(7):
        valid = ci[q] >= 0
        key = np.where(valid, -cd[q] if ip else cd[q], np.inf)
        order = np.lexsort((ci[q], key))[:kk]
        assert (ci[q][order] == If[q, :kk]).all()


@settings(max_examples=20, deadline=None)
@given(n=st.integers(30, 500), seed=st.integers(0, 10_000))
def test_blas_and_direct_paths_agree_within_fp32(oracle, n, seed):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, 24)).astype(np.float32)
    xq = rng.standard_normal((5, 24)).astype(np.float32)
    idx = oracle.FlatIndexOracle(24, oracle.METRIC_L2)
    idx.add(xb)
    Da, Ia = idx.search(xq, 5, direct=False)
    Db, Ib = idx.search(xq, 5, direct=True)
    st_ = oracle.compare_topk(Da, Ia, Db, Ib, lambda ids: idx.exact_scores(xq, ids), oracle.METRIC_L2, 1e-5, 1e-4)
    assert st_["recall"] >= 0.99


def test_oracle_agrees_with_sklearn_bruteforce(oracle):
    """FAISS (the reference's dependency) is not installable here, so as an INDEPENDENT third-party check of the
    oracle's flat-kNN semantics: scikit-learn's brute-force NearestNeighbors must return the same neighbours
    (squared-L2 ascending == euclidean ascending; inner product on normalised rows == cosine similarity descending)."""
    import numpy as np
    sk = pytest.importorskip("sklearn.neighbors")
    rng = np.random.default_rng(2024)
    xb = rng.standard_normal((4000, 48)).astype(np.float32)
    xq = rng.standard_normal((60, 48)).astype(np.float32)
    k = 12
    ref = oracle.FlatIndexOracle(48, oracle.METRIC_L2)
    ref.add(xb)
    D, I = ref.search(xq, k, direct=True)
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(xb.astype(np.float64))
    Ds, Is = nn.kneighbors(xq.astype(np.float64))
    np.testing.assert_array_equal(I, Is)
    np.testing.assert_allclose(D, Ds ** 2, rtol=2e-5, atol=1e-4)
    xbn, xqn = oracle.maybe_normalize(xb, True), oracle.maybe_normalize(xq, True)
    refc = oracle.FlatIndexOracle(48, oracle.METRIC_IP)
    refc.add(xbn)
    Dc, Ic = refc.search(xqn, k)
    nnc = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(xb.astype(np.float64))
    Dsc, Isc = nnc.kneighbors(xq.astype(np.float64))
    np.testing.assert_array_equal(Ic, Isc)
    np.testing.assert_allclose(Dc, 1.0 - Dsc, rtol=0, atol=2e-6)
