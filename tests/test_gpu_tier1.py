"""GPU parity for the tiered certified search of fp32 stores ('Split-precision mode'):
tier 1 = ONE tensor-core term on the bf16 roundings + 128 candidates + exact fp32 re-rank + certificate against the bf16
error bound (k <= 64, >= 262144 rows); queries it cannot certify are compacted and go through the three-term pass
(tier 2); what that cannot certify goes to the exact CUDA-core kernel.  Whatever tier certifies a query, the result
must be the exact-fp32 neighbours (north star: 1e-5 relative for fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_F32 = 1e-5


def _gauss(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def _check_vs_oracle(pkg, oracle, idx, xb, xq, D, I, k, metric, cos):
    ref = oracle.FlatIndexOracle(xb.shape[1], metric)
    ref.add(oracle.maybe_normalize(xb, cos))
    qn = oracle.maybe_normalize(xq, cos)
    Dr, Ir = ref.search(qn, k + 8, direct=False)
    if metric == pkg.METRIC_L2:
        floor = 2e-6 * float((qn * qn).sum(1).max() + (ref._base() ** 2).sum(1).max())
    else:
        floor = 1e-6
    st = oracle.compare_topk(D, I, Dr, Ir, lambda ids: ref.exact_scores(qn, ids), metric, tol=TOL_F32, abs_floor=floor)
    assert st["recall"] == 1.0, st


@pytest.mark.parametrize("metric_s,cos,k", [("IP", True, 10), ("L2", False, 15), ("IP", False, 32), ("L2", False, 64)])
def test_tier1_certifies_gaussian(pkg, oracle, metric_s, cos, k):
    """Well-separated data: tier 1 certifies (nearly) everything; neighbours == oracle; identical to the search with
    tier 1 switched off (both end in the same exact fp32 re-rank, so distances agree bit-for-bit)."""
    N, Dm, Q = 300_000, 64, 300
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    xb, xq = _gauss(N, Dm, 1234), _gauss(Q, Dm, 5678)
    xq[::7] = xb[: len(xq[::7])] + 0.05 * _gauss(len(xq[::7]), Dm, 9)
    xq[1] = xb[5]
    idx = pkg.FlatIndex(Dm, metric, "f32")
    for s in range(0, N, 65536):
        idx.add(xb[s:s + 65536], normalize=cos)
    D, I = idx.search(xq, k, normalize=cos)
    t1q, t1u = idx.last_tier1
    assert t1q == Q and t1u <= Q // 10, (t1q, t1u)
    assert idx.last_tier1_candidates == (32 if k <= 16 else 128)
    assert idx.last_kernel_ms()[1] == "tc"
    _check_vs_oracle(pkg, oracle, idx, xb, xq, D, I, k, metric, cos)
    idx.set_option("tier1", 0)
    D3, I3 = idx.search(xq, k, normalize=cos)
    assert idx.last_tier1 == (0, 0)
    np.testing.assert_array_equal(I, I3)
    np.testing.assert_array_equal(D, D3)


def test_tier_chain_when_nothing_certifies(pkg, oracle):
    """Every query sits on a block of 201 identical lattice rows: the 128 candidates tie the k-th neighbour, so neither
    tensor-core tier can certify and every query walks tier 1 -> tier 2 -> exact kernel; the result must still equal
    the oracle bit-for-bit (lowest ids of the block first).  The next searches skip tier 1 (it failed for most queries)
    and must return the same."""
    rng = np.random.default_rng(3)
    N, Dm, Q, k = 270_000, 32, 40, 10
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    for j in range(5):
        xb[10_000 * (j + 1):10_000 * (j + 1) + 200] = xb[j]
    xq = np.stack([xb[i % 5] for i in range(Q)])
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    assert Ir[3, 0] == 3 and Ir[3, 1] == 40_000
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "f32")
    idx.add(xb)
    D, I = idx.search(xq, k)
    t1q, t1u = idx.last_tier1
    assert t1q == Q and t1u == Q, (t1q, t1u)
    assert idx.last_tier1_candidates == 32 and idx.last_uncertified == Q
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    D2, I2 = idx.search(xq, k)                         # 32 candidates failed for > 25 %: 128 candidates now
    assert idx.last_tier1 == (Q, Q) and idx.last_tier1_candidates == 128
    np.testing.assert_array_equal(I2, Ir)
    np.testing.assert_array_equal(D2, Dr)
    D3, I3 = idx.search(xq, k)                         # 128 failed for > 50 %: tier 1 is skipped for a while
    assert idx.last_tier1 == (0, 0) and idx.last_tier1_candidates == 0
    np.testing.assert_array_equal(I3, Ir)
    np.testing.assert_array_equal(D3, Dr)


def test_lattice_with_ties_inside_the_candidates(pkg, oracle):
    """Plain lattice data: ties at the k-th boundary fall INSIDE the 128 candidates, where the exact re-rank orders them
    (lowest id first); integer keys clear the bf16 bound, so tier 1 certifies -- and must equal the oracle bit-for-bit."""
    rng = np.random.default_rng(4)
    N, Dm, Q, k = 270_000, 32, 40, 10
    xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
    xq = rng.integers(-2, 3, size=(Q, Dm)).astype(np.float32)
    ref = oracle.FlatIndexOracle(Dm, pkg.METRIC_L2)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, direct=False)
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "f32")
    idx.add(xb)
    D, I = idx.search(xq, k)
    assert idx.last_tier1[0] == Q
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)


def test_tier1_partial_certification_clustered(pkg, oracle):
    """Clustered rows with duplicated members: queries next to a block of > 128 identical rows cannot be certified by
    tier 1 (nor tier 2) while the rest can; every query must come back exact, duplicates in ascending id order."""
    N, Dm, Q, k = 280_000, 48, 200, 10
    xb = _gauss(N, Dm, 21)
    xb[100_000:100_200] = xb[100_000]                  # 200 identical rows
    xq = _gauss(Q, Dm, 22)
    xq[4] = xb[100_000]
    xq[9] = xb[100_000] + 1e-4
    labels = (np.arange(N) % 2).astype(np.float32)
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "f32")
    idx.add(xb)
    idx.set_labels(labels)
    D, I, L = idx.search(xq, k, return_labels=True)
    t1q, t1u = idx.last_tier1
    assert t1q == Q and 2 <= t1u <= Q // 4, (t1q, t1u)
    assert idx.last_uncertified >= 2
    np.testing.assert_array_equal(I[4], np.arange(100_000, 100_010))
    np.testing.assert_array_equal(L, labels[I])
    _check_vs_oracle(pkg, oracle, idx, xb, xq, D, I, k, pkg.METRIC_L2, False)
    Ds, Is = idx.search(xq, k, algo="simt")
    np.testing.assert_array_equal(I, Is)
    np.testing.assert_allclose(D, Ds, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("scenario", ["clustered", "nothing_certifies", "gaussian_ip"])
def test_device_path_is_stream_ordered(pkg, oracle, scenario):
    """Device-tensor searches of fp32 stores never block the host: how many queries each tier must re-search is known on
    the device only, and every follow-up launch (compaction, three-term pass, exact CUDA-core search, scatter) is sized
    there from the count (DevPlan) instead of reading it back.  Same answers as the host path, zero host
    synchronisations inside `search`, also when every query walks all three tiers."""
    import torch
    if scenario == "nothing_certifies":
        rng = np.random.default_rng(3)
        N, Dm, Q, k, metric = 270_000, 32, 300, 10, pkg.METRIC_L2
        xb = rng.integers(-2, 3, size=(N, Dm)).astype(np.float32)
        for j in range(5):
            xb[10_000 * (j + 1):10_000 * (j + 1) + 200] = xb[j]
        xq = np.stack([xb[i % 5] for i in range(Q)])
    elif scenario == "clustered":
        N, Dm, Q, k, metric = 280_000, 48, 700, 15, pkg.METRIC_L2
        xb = _gauss(N, Dm, 21)
        xb[100_000:100_200] = xb[100_000]
        xq = _gauss(Q, Dm, 22)
        xq[4] = xb[100_000]
        xq[9] = xb[100_000] + 1e-4
        xq[650] = xb[100_007]
    else:
        N, Dm, Q, k, metric = 300_000, 96, 1000, 10, pkg.METRIC_IP
        xb, xq = _gauss(N, Dm, 31), _gauss(Q, Dm, 32)
    idx = pkg.FlatIndex(Dm, metric, "f32")
    idx.add(xb)
    Dh, Ih = idx.search(xq, k)                          # host path (synchronises by contract)
    uncert_host = idx.last_uncertified
    t1_host = idx.last_tier1
    idx2 = pkg.FlatIndex(Dm, metric, "f32")             # fresh adaptive state: same tier decisions as the host run
    idx2.add(xb)
    q = torch.from_numpy(xq).cuda()
    warm = pkg.FlatIndex(Dm, metric, "f32")
    warm.add(xb[:270_000])
    warm.search(q, k)                                   # (context / module warm-up only)
    s0 = idx2.host_sync_count
    Dd, Id = idx2.search(q, k)
    assert idx2.host_sync_count == s0, "the device-tensor search blocked the host"
    torch.cuda.synchronize()
    np.testing.assert_array_equal(Id.cpu().numpy(), Ih)
    np.testing.assert_array_equal(Dd.cpu().numpy(), Dh)
    assert idx2.last_tier1 == t1_host and idx2.last_uncertified == uncert_host      # counters arrive behind the batch
    if scenario == "nothing_certifies":
        assert t1_host == (Q, Q) and uncert_host == Q
    if scenario == "clustered":
        assert uncert_host >= 3



def test_two_list_cover_equals_32_entry_lists_and_flags_saturated_lists(pkg, oracle):
    """Tier 1 keeps its 32 candidates as a two-list cover of 16-entry lists (SelectSmall<16, 2>).  (a) Same answers as
    with 32-entry lists (option tier1_share2 = 0) and as the oracle.  (b) A query whose 40 nearest rows sit in ONE
    candidate list (contiguous rows of one half-tile), spaced 0.004 apart in exact distance -- far above the fp32
    tolerance, far below the bf16 error of the one-term keys -- saturates that list: the 16 entries it keeps are the
    best by APPROXIMATE key, so the true top-10 is partly missing and the query must not be certified by tier 1 (the
    merge flags it; tier 2 then finds the exact neighbours)."""
    N, Dm, Q, k = 300_000, 64, 300, 10
    rng = np.random.default_rng(21)
    xb, xq = _gauss(N, Dm, 31), _gauss(Q, Dm, 32)
    hard = list(range(0, 16, 2))
    for t, qi in enumerate(hard):
        r0 = 1024 * (3 + 7 * t)                                   # start of a database tile -> one 128-row half
        u = rng.standard_normal((40, Dm)).astype(np.float32)
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        xb[r0:r0 + 40] = xq[qi] + np.sqrt(0.004 * np.arange(40, dtype=np.float32))[:, None] * u
    idx = pkg.FlatIndex(Dm, pkg.METRIC_L2, "f32")
    for s in range(0, N, 65536):
        idx.add(xb[s:s + 65536])
    D, I = idx.search(xq, k)
    t1q, t1u = idx.last_tier1
    assert t1q == Q and idx.last_tier1_candidates == 32
    assert t1u >= len(hard), (t1q, t1u)                           # every saturated query went on to tier 2
    for t, qi in enumerate(hard):
        r0 = 1024 * (3 + 7 * t)
        np.testing.assert_array_equal(I[qi], np.arange(r0, r0 + k))
    _check_vs_oracle(pkg, oracle, idx, xb, xq, D, I, k, pkg.METRIC_L2, False)
    idx.set_option("tier1_share2", 0)
    D0, I0 = idx.search(xq, k)
    np.testing.assert_array_equal(I, I0)
    np.testing.assert_array_equal(D, D0)
    idx.close()
