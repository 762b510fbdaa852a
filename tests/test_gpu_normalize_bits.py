"""Cosine mode stores `arr / (np.linalg.norm(arr, axis=1, keepdims=True) + 1e-12)` (reference vector_database.py:100-105,
computed by numpy on the host).  The ingest / query-prep kernels reproduce numpy's pairwise summation order, so the
stored rows -- what `index.reconstruct` hands to RADADModel -- and the normalised queries are BIT-IDENTICAL to the
reference's, not merely within rounding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIMS = [1, 3, 7, 8, 9, 15, 56, 96, 100, 128, 130, 255, 256, 448, 768, 1000, 1020, 1024, 1030, 2000, 5376, 9000]


def _ref_normalize(x):
    return x / (np.linalg.norm(x, axis=1, keepdims=True) + 1e-12)        # vector_database.py:102-104, verbatim


@pytest.mark.parametrize("d", DIMS)
def test_stored_rows_bit_equal_numpy(pkg, d):
    rng = np.random.default_rng(d)
    x = (rng.standard_normal((301, d)) * rng.uniform(1e-3, 50.0, size=(301, 1))).astype(np.float32)
    x[5] = 0.0                                                            # zero row stays zero (0 / 1e-12)
    x[6] = 1e-20                                                          # squares underflow: norm = 0 in numpy too
    x[7] *= 1e-25                                                         # denormal squares / tiny quotients
    x[8] *= 1e17                                                          # squares near overflow
    x[9, ::3] = 0.0                                                       # sparse row (ReLU-like exact zeros)
    x[10, ::2] = -0.0                                                     # negative zeros keep their sign
    x[11, 0] = 3e38 if d > 1 else x[11, 0]                                # |x|^2 overflows: norm = inf, row -> 0
    x[12] = np.where(np.arange(d) % 5 == 0, x[12] * 1e-30, x[12])         # tiny components next to normal ones
    idx = pkg.FlatIndex(d, pkg.METRIC_IP, "f32")
    idx.add(x[:100], normalize=True)
    import torch
    idx.add(torch.from_numpy(x[100:]).cuda(), normalize=True)            # device-tensor ingest takes the same kernel
    got = idx.reconstruct_batch(np.arange(301))
    ref = _ref_normalize(x)
    np.testing.assert_array_equal(got.view(np.uint32), ref.astype(np.float32).view(np.uint32))


@pytest.mark.parametrize("algo,nq", [("stream", 1), ("stream", 3), ("simt", 40), ("tc", 200)])
@pytest.mark.parametrize("d", [96, 768, 1000])
def test_normalised_queries_bit_equal_numpy(pkg, oracle, algo, nq, d):
    """Database = the d unit vectors (+ zero rows up to 300): <q_hat, e_i> = q_hat[i] exactly, so the inner products the
    search returns ARE the normalised query's largest components -- compared bit for bit with numpy's."""
    eye = np.zeros((max(d, 300), d), dtype=np.float32)
    eye[:d] = np.eye(d, dtype=np.float32)
    store = "bf16" if algo == "tc" else "f32"
    idx = pkg.FlatIndex(d, pkg.METRIC_IP, store)
    idx.add(eye)
    rng = np.random.default_rng(7 * d + nq)
    q = (rng.standard_normal((nq, d)) * 3.0).astype(np.float32)
    k = 8
    D, I = idx.search(q, k, normalize=True, algo=algo)
    qn = _ref_normalize(q).astype(np.float32)
    if store == "bf16":
        qn = oracle.round_bf16(qn)                                       # the 16-bit scorer sees the rounded query
    want = -np.sort(-qn, axis=1)[:, :k]
    np.testing.assert_array_equal(D.view(np.uint32), want.view(np.uint32))
    np.testing.assert_array_equal(I, np.argsort(-qn, axis=1, kind="stable")[:, :k])
