"""GPU parity for the pipelined host-buffer search (`search_impl` in csrc/radad_flat.cu): a search with HOST buffers of
>= 4096 queries / 8 MB uploads the batch in pieces on a second stream so that the copy of piece i + 1 overlaps the search
of piece i (the call the reference makes: `vector_database.py:152-190` hands `index.search` a numpy batch).  The cut
must be invisible: identical ids / distances / labels to the unpipelined call and to the device-tensor call, and the
oracle's neighbours on lattice data (bit-exact), for both piece schedules (compute-bound: growing pieces, copy-bound:
equal pieces), 16-bit and fp32 stores, ragged sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lattice(n, d, seed):
    return np.random.default_rng(seed).integers(-3, 4, size=(n, d)).astype(np.float32)


@pytest.mark.parametrize("store,metric_s,n,nq,d,k", [
    ("bf16", "L2", 5000, 9001, 256, 10),      # copy-bound schedule (equal pieces), ragged tail
    ("bf16", "IP", 401_000, 8192, 256, 15),   # compute-bound schedule (growing pieces)
    ("f32", "L2", 6000, 8200, 256, 10),       # fp32 store, small shard: split-precision certified path per piece
    ("f32", "IP", 401_000, 8200, 256, 10),    # fp32 store, tiered certified search per piece
])
def test_pipelined_host_search_equals_unpipelined(pkg, oracle, store, metric_s, n, nq, d, k):
    import torch
    metric = pkg.METRIC_IP if metric_s == "IP" else pkg.METRIC_L2
    xb, xq = _lattice(n, d, 11), _lattice(nq, d, 12)
    xq[:64] = xb[:64]
    labels = (np.arange(n) % 2).astype(np.float32)
    idx = pkg.FlatIndex(d, metric, store, device=0)
    idx.add(xb)
    idx.set_labels(labels)
    assert xq.nbytes >= 8 << 20
    D1, I1 = idx.search(xq, k)                       # pipelined (default)
    syncs_piped = idx.host_sync_count
    idx.set_option("host_pipeline", 0)
    D0, I0 = idx.search(xq, k)                       # one upload, then the search
    np.testing.assert_array_equal(I1, I0)
    np.testing.assert_array_equal(D1, D0)
    Dt, It = idx.search(torch.from_numpy(xq).cuda(), k)
    np.testing.assert_array_equal(I1, It.cpu().numpy())
    np.testing.assert_array_equal(D1, Dt.cpu().numpy())
    # oracle on a slice of the queries that straddles the first piece border of either schedule
    sel = np.r_[0:96, 1000:1064, 2040:2120, nq - 70:nq]
    ref = oracle.FlatIndexOracle(d, metric)
    ref.add(xb)
    Dr, Ir = ref.search(xq[sel], k)
    np.testing.assert_array_equal(I1[sel], Ir)
    np.testing.assert_array_equal(D1[sel], Dr)
    # the pipelined call blocks the host once per batch, like the plain one
    assert syncs_piped >= 1
    idx.close()


def test_pipelined_host_search_through_vector_database(pkg, make_cfg):
    """Same through the reference-facing wrapper (search_batch with a pageable numpy batch, cosine)."""
    rng = np.random.default_rng(5)
    n, nq, d = 30_000, 4500, 512
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    cfg = make_cfg("IP", db_dtype="bf16")
    db = pkg.VectorDatabase(cfg)
    db.create_index(d)
    db.add_vectors(xb, [f"p{i}" for i in range(n)], [int(i % 2) for i in range(n)], {"split": "train"})
    D1, I1 = db.search_batch(xq, 10)[:2]
    db.index.set_option("host_pipeline", 0)
    D0, I0 = db.search_batch(xq, 10)[:2]
    np.testing.assert_array_equal(np.asarray(I1), np.asarray(I0))
    np.testing.assert_array_equal(np.asarray(D1), np.asarray(D0))


def test_copy_async_moves_bytes_in_stream_order(pkg):
    """rdb_copy_async (the exchange primitive of the single-process multi-GPU search): a copy kernel on the index's own
    GPU and the caller's stream -- 16-byte path and the byte path (odd sizes / unaligned views), ordered behind the
    producer of the source.  The peer-memory case is covered by the multi-GPU tests."""
    import torch
    idx = pkg.FlatIndex(8, pkg.METRIC_L2, "f32", device=0)
    g = torch.Generator(device="cuda:0")
    g.manual_seed(3)
    src = torch.randn((1000, 257), generator=g, device="cuda:0")
    dst = torch.zeros_like(src)
    idx.copy_async(dst, src * 2.0)                       # ordered behind the multiply on the same stream
    assert torch.equal(dst, src * 2.0)
    a = torch.arange(10_001, dtype=torch.uint8, device="cuda:0")
    b = torch.zeros(10_001, dtype=torch.uint8, device="cuda:0")
    idx.copy_async(b[1:], a[1:])                         # unaligned, odd length -> byte path
    assert torch.equal(b[1:], a[1:]) and int(b[0]) == 0
    with pytest.raises(RuntimeError):
        idx.copy_async(dst[:10], src[:11])
    idx.close()
