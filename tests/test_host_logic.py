"""CPU: host-side logic that needs no GPU -- shard partitioning, candidate packing, comparator, the
device-resident retrieval's basename coding."""
import numpy as np
import pytest
import torch


def test_shard_bounds_cover_and_order(pkg):
    for n in (0, 1, 7, 256, 10_000_000, 99_999_999):
        for g in (1, 2, 4, 8):
            spans = [pkg.shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            assert max(e - s for s, e in spans) == -(-n // g)


def test_pack_unpack_roundtrip(pkg):
    from importlib import import_module
    sh = import_module(pkg.__name__ + ".sharded")
    g = torch.Generator().manual_seed(0)
    key = torch.randn(5, 7, generator=g)
    key[0, 0] = float("-inf")
    gid = torch.randint(-1, 2**40, (5, 7), generator=g, dtype=torch.int64)
    lab = torch.randint(0, 2, (5, 7), generator=g).float()
    p = sh.pack_candidates(key, gid, lab)
    assert p.dtype == torch.int32 and p.shape == (5, 7, 4)
    k2, g2, l2 = sh.unpack_candidates(p)
    assert torch.equal(k2, key) and torch.equal(g2, gid) and torch.equal(l2, lab)


def test_comparator_accepts_ties_and_rejects_errors(oracle):
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((500, 16)).astype(np.float32)
    xb[100] = xb[7]                       # exact duplicate -> tie
    q = xb[[7, 20]].copy()
    idx = oracle.FlatIndexOracle(16, oracle.METRIC_L2)
    idx.add(xb)
    D, I = idx.search(q, 13)
    exact = lambda ids: idx.exact_scores(q, ids)   # noqa: E731
    st = oracle.compare_topk(D[:, :5], I[:, :5], D, I, exact, oracle.METRIC_L2, 1e-5)
    assert st["recall"] == 1.0
    swapped = I[:, :5].copy()
    swapped[0, [0, 1]] = swapped[0, [1, 0]]        # ids 7 and 100 tie at distance 0: either order is fine
    oracle.compare_topk(D[:, :5], swapped, D, I, exact, oracle.METRIC_L2, 1e-5)
    wrong = I[:, :5].copy()
    wrong[1, 4] = 499 if 499 not in I[1] else 498
    with pytest.raises(AssertionError):
        oracle.compare_topk(D[:, :5], wrong, D, I, exact, oracle.METRIC_L2, 1e-5)


def test_basename_codes(pkg):
    from importlib import import_module
    r = import_module(pkg.__name__ + ".retrieval")

    class V:
        vector_paths = ["/a/x.wav", "/b/x.wav", "/a/y.wav"]
    codes, table = r._basename_codes(V)
    assert codes.tolist() == [0, 0, 1] and table == {"x.wav": 0, "y.wav": 1}


def test_multi_gpu_water_filling_plan(pkg):
    """Placement of add() calls over the shards of MultiGpuFlatIndex (pure host logic; no GPU needed)."""
    m = pkg.MultiGpuFlatIndex.__new__(pkg.MultiGpuFlatIndex)

    class _S:
        def __init__(self):
            self.ntotal = 0
    m.shards = [_S() for _ in range(4)]

    def apply(n):
        plan = m._plan(n)
        assert sum(c for _, c in plan) == n and all(c > 0 for _, c in plan)
        for g, c in plan:
            m.shards[g].ntotal += c
        return plan

    assert apply(100) == [(0, 100)]                        # tiny add: whole call to the least-loaded shard
    assert apply(100) == [(1, 100)]
    apply(1_000_000)
    sizes = [s.ntotal for s in m.shards]
    assert sum(sizes) == 1_000_200 and max(sizes) - min(sizes) <= 1
    apply(10_000)                                          # reference slice size (vector_add_batch_size)
    sizes = [s.ntotal for s in m.shards]
    assert sum(sizes) == 1_010_200 and max(sizes) - min(sizes) <= 1


def test_lockstep_membership_arithmetic():
    """Mirror of the lock-step window bookkeeping (csrc/score_tc.cuh producer + launch_tc in csrc/radad_flat.cu):
    every work unit must fall into exactly one (slot, chunk-in-slot) counter group, the chunk-in-slot index must stay
    below the `span` the host allocates for, and the `members` count every producer derives for its group must equal
    the number of units that really land in it -- otherwise waiters would expect arrivals that never come (and give
    the window up) or a group would be released early."""
    import random
    rnd = random.Random(7)
    cases = [(512, 13, 148), (256, 13, 74), (128, 37, 148), (128, 8, 148), (64, 8, 74), (40, 24, 74), (75, 5, 148)]
    cases += [(rnd.randint(1, 600), rnd.randint(1, 40), rnd.choice([74, 148, 7, 31])) for _ in range(200)]
    for nqt, S, slots_avail in cases:
        num_units = nqt * S
        ngroups = min(num_units, slots_avail)
        span = (ngroups + nqt - 1) // nqt + 1                     # host: chunks one slot can touch
        seen = {}
        for group in range(ngroups):                              # device: persistent loop of one CTA (pair)
            slot = 0
            for unit in range(group, num_units, ngroups):
                chunk = unit // nqt
                u_lo = max(slot * ngroups, chunk * nqt)
                u_hi = min((slot + 1) * ngroups, (chunk + 1) * nqt, num_units)
                members = u_hi - u_lo
                cidx = chunk - (slot * ngroups) // nqt
                assert 0 <= cidx < span, (nqt, S, ngroups, unit, cidx, span)
                assert u_lo <= unit < u_hi
                seen.setdefault((slot, cidx), []).append(members)
                slot += 1
        total = 0
        for key, ms in seen.items():
            assert len(set(ms)) == 1 and ms[0] == len(ms), (nqt, S, ngroups, key, ms)
            total += len(ms)
        assert total == num_units


def test_reconstruct_cache_logic_without_gpu(pkg):
    """flat_index._ReconstructCache (host logic only): lazy -- nothing is fetched until reconstruct is called; ONE
    batched fetch of the unique valid ids of the last search; hits are independent copies; misses, ids of an older
    search and oversized batches fall through to the single-row path."""
    import numpy as np
    from importlib import import_module
    fi = import_module(pkg.__name__ + ".flat_index")

    class Fake(fi._ReconstructCache):
        d = 4

        def __init__(self):
            self.batched, self.single = [], []

        def reconstruct_batch(self, ids):
            self.batched.append(np.array(ids))
            return np.stack([np.full(4, float(i), np.float32) for i in ids])

        def reconstruct(self, i):
            hit = self._rc_lookup(int(i))
            if hit is not None:
                return hit
            self.single.append(int(i))
            return np.full(4, float(i), np.float32)

    f = Fake()
    assert f.reconstruct(3)[0] == 3 and f.single == [3] and not f.batched        # no search yet: single-row path
    f._rc_note_search(np.array([[5, 2, -1], [2, 9, 5]], dtype=np.int64))
    assert not f.batched                                                            # lazy
    a = f.reconstruct(9)
    assert len(f.batched) == 1 and list(f.batched[0]) == [2, 5, 9]                  # unique, valid, sorted
    a[:] = -1
    assert f.reconstruct(9)[0] == 9 and f.reconstruct(2)[0] == 2 and len(f.batched) == 1
    assert f.reconstruct(7)[0] == 7 and f.single == [3, 7]                          # not in the result: miss
    f._rc_note_search(np.array([[1]], dtype=np.int64))                              # next search drops the cache
    assert f.reconstruct(5)[0] == 5 and f.single == [3, 7, 5] and len(f.batched) == 2
    f._RC_MAX_BYTES = 8                                                             # oversized batch: never fetched
    f._rc_note_search(np.array([[4, 6]], dtype=np.int64))
    assert f.reconstruct(4)[0] == 4 and len(f.batched) == 2 and f.single[-1] == 4
