"""CPU mirror of the warp-level list folding used by the streaming scorer and the mid-batch merge
(csrc/merge.cuh: merge32 / sort32 / block_tree_merge32): 32 lanes are a numpy vector, a shuffle is an index
permutation.  Checks the network itself -- best 32 of the union of two sorted lists, sorted -- incl. empty slots (0),
duplicates of the empty word and full ties on the key part; the CUDA code is exercised against the oracle by the GPU
tests (tests/test_gpu_parity.py, tests/test_gpu_merge_tree.py)."""
import numpy as np

LANES = np.arange(32)


def merge32(a, b):
    br = b[31 - LANES]
    m = np.maximum(a, br)
    o = 16
    while o > 0:
        x = m[LANES ^ o]
        keep_small = (LANES & o) != 0
        m = np.where((m < x) == keep_small, m, x)
        o >>= 1
    return m


def sort32(m):
    size = 2
    while size <= 32:
        o = size >> 1
        while o > 0:
            x = m[LANES ^ o]
            desc = (LANES & size) == 0
            keep_small = ((LANES & o) != 0) == desc
            m = np.where((m < x) == keep_small, m, x)
            o >>= 1
        size <<= 1
    return m


def tree(lists):
    """block_tree_merge32 over a power-of-two number of warp-held lists."""
    lists = list(lists)
    s = 1
    while s < len(lists):
        for w in range(0, len(lists), 2 * s):
            lists[w] = merge32(lists[w], lists[w + s])
        s <<= 1
    return lists[0]


def _sorted_list(rng, n_valid, key_bits):
    keys = rng.integers(1, 1 << key_bits, size=n_valid, dtype=np.uint64)
    ids = rng.permutation(1 << 20)[:n_valid].astype(np.uint64)
    w = (keys << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - ids)
    w = np.sort(w)[::-1]
    return np.concatenate([w, np.zeros(32 - n_valid, np.uint64)])


def test_merge32_is_best32_of_union():
    rng = np.random.default_rng(0)
    for trial in range(300):
        kb = [3, 8, 30][trial % 3]                      # 3 bits: massive ties on the key part
        a = _sorted_list(rng, int(rng.integers(0, 33)), kb)
        b = _sorted_list(rng, int(rng.integers(0, 33)), kb)
        want = np.sort(np.concatenate([a, b]))[::-1][:32]
        np.testing.assert_array_equal(merge32(a, b), want)


def test_sort32_sorts_descending():
    rng = np.random.default_rng(1)
    for _ in range(200):
        m = rng.integers(0, 1 << 40, size=32, dtype=np.uint64)
        m[rng.integers(0, 32, size=5)] = 0
        np.testing.assert_array_equal(sort32(m.copy()), np.sort(m)[::-1])


def test_tree_of_16_lists():
    rng = np.random.default_rng(2)
    for _ in range(50):
        lists = [_sorted_list(rng, int(rng.integers(0, 33)), 10) for _ in range(16)]
        want = np.sort(np.concatenate(lists))[::-1][:32]
        np.testing.assert_array_equal(tree(lists), want)
