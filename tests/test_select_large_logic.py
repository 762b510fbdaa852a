"""CPU mirror of the radix select in csrc/select_large.cuh (the k > 128 path): same packing, same 8-bit MSB-first passes,
same early exit, checked against a plain sort.  Guards the bookkeeping (bin walk, 'whole bin wanted' exit, threshold
semantics) without a GPU; the kernel itself is parity-tested in tests/test_gpu_large_k.py."""
import numpy as np
import pytest


def _ordered_f32(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    neg = (u & 0x80000000) != 0
    return np.where(neg, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def _pack(keys, row0):
    rows = np.arange(row0, row0 + len(keys), dtype=np.uint64)
    return (_ordered_f32(keys) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - rows)


def radix_select_mirror(keys, row0, k):
    """Returns (selected packed values sorted descending, number of passes) exactly as the kernel computes them."""
    v = _pack(keys, row0)
    kk = min(k, len(v))
    prefix, mask, passes = np.uint64(0), np.uint64(0), 0
    if len(v) > k:
        remaining, done = kk, False
        for p in range(8):
            if done:
                break
            passes += 1
            shift = np.uint64(56 - 8 * p)
            match = (v & mask) == prefix
            hist = np.bincount(((v[match] >> shift) & np.uint64(255)).astype(np.int64), minlength=256)
            cum, b = 0, 255
            while b > 0:
                if cum + hist[b] >= remaining:
                    break
                cum += hist[b]
                b -= 1
            done = hist[b] == remaining - cum
            remaining -= cum
            prefix |= np.uint64(b) << shift
            mask |= np.uint64(0xFF) << shift
    sel = v[v >= prefix]
    assert len(sel) == kk
    return np.sort(sel)[::-1], passes


@pytest.mark.parametrize("seed", range(6))
def test_radix_select_mirror_matches_sort(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5000))
    k = int(rng.integers(129, 2049))
    kind = seed % 3
    if kind == 0:
        keys = rng.standard_normal(n).astype(np.float32)
    elif kind == 1:
        keys = rng.integers(-3, 4, n).astype(np.float32)              # lattice: few distinct keys, ties at the boundary
    else:
        keys = np.full(n, 16.0, np.float32)                           # all equal: every digit down to the id decides
    row0 = int(rng.integers(0, 1 << 20))
    got, passes = radix_select_mirror(keys, row0, k)
    want = np.sort(_pack(keys, row0))[::-1][:min(k, n)]
    np.testing.assert_array_equal(got, want)
    # decoded order = (key descending, row id ascending)
    rows = (np.uint64(0xFFFFFFFF) - (got & np.uint64(0xFFFFFFFF))).astype(np.int64)
    order = np.lexsort((np.arange(row0, row0 + n), -keys.astype(np.float64)))[:min(k, n)] + row0
    np.testing.assert_array_equal(rows, order)
    if kind == 2 and n > k:
        assert passes >= 5


def sampled_select_mirror(keys, row0, k, cap=8192, sample=16384, min_len=32768):
    """Fast path of select_dense_kernel: pivot from a strided sample, one collect pass, exact iff k <= count <= cap.
    Returns (selected or None when the kernel would fall back, count)."""
    n = len(keys)
    kk = min(k, n)
    if n < min_len or n <= k:
        return None, 0
    st = n // sample
    ns = (n + st - 1) // st
    r = max(32, (5 * kk // 2 + st - 1) // st)
    if r >= ns:
        return None, 0
    v = _pack(keys, row0)
    vs = v[::st]
    assert len(vs) == ns
    T = np.sort(vs)[::-1][r - 1]                       # what the radix select on the sample returns
    sel = v[v >= T]
    if kk <= len(sel) <= cap:
        return np.sort(sel)[::-1][:kk], len(sel)
    return None, len(sel)


@pytest.mark.parametrize("n,k", [(40_000, 129), (300_000, 1000), (1_000_000, 2048), (70_001, 500)])
def test_sampled_pivot_fast_path_is_exact_and_usually_taken(n, k):
    keys = np.random.default_rng(n + k).standard_normal(n).astype(np.float32)
    got, count = sampled_select_mirror(keys, 12345, k)
    assert got is not None, count                      # Gaussian keys: the pivot lands between k and the buffer size
    assert k <= count <= 8192
    np.testing.assert_array_equal(got, np.sort(_pack(keys, 12345))[::-1][:k])


def test_sampled_pivot_with_heavy_ties():
    """(key, id) pairs are distinct, so ties on the key do not pile up at the pivot: the fast path stays exact and is
    still taken (ids order the ~28 K elements per key value)."""
    keys = np.random.default_rng(5).integers(-3, 4, 200_000).astype(np.float32)
    got, count = sampled_select_mirror(keys, 0, 300)
    assert got is not None and 300 <= count <= 8192
    np.testing.assert_array_equal(got, np.sort(_pack(keys, 0))[::-1][:300])


def test_sampled_pivot_falls_back_when_the_sample_misleads():
    """Adversarial layout: the large keys sit exactly on the sample stride, so the pivot is far too high for the row
    -> fewer than k elements reach it -> the kernel takes the exact path."""
    n, k = 163_840, 1000
    keys = np.zeros(n, np.float32)
    keys[::10] = 5.0                                   # stride = n // 16384 = 10: every sampled element is a 5.0
    keys[1::10] = np.linspace(1.0, 4.0, len(keys[1::10]), dtype=np.float32)
    got, count = sampled_select_mirror(keys, 0, k)
    assert got is None and count < k
