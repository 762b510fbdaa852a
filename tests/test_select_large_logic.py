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
