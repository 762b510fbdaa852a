"""``FlatIndex`` -- the object behind ``VectorDatabase.index``.

Duck type of the faiss flat index the reference reaches through ``self.index``
(``vector_database.py:138,151,169,181,210``; ``pipeline.py:465,503,1039``; ``app.py:75,246``):
``add(x)``, ``search(q, k) -> (D, I)``, ``reconstruct(i)``, ``ntotal``, ``d``, ``is_trained``,
``train(x)``, ``nprobe``.  All arithmetic runs in ``libradad_flat.so`` (hand-written sm_100a CUDA)
through the C ABI of ``include/radad_flat.h``; numpy arrays are accepted at the reference boundary,
torch CUDA tensors on the device-resident fast path (no host round trip).
"""
from __future__ import annotations

import ctypes
import functools
import threading
from typing import Optional, Tuple

import numpy as np

if __package__:
    from . import _cabi
else:           # flat layout: this directory itself is on sys.path, as the reference's `from vector_database import ...`
    import _cabi
(ALGO_AUTO, ALGO_SIMT, ALGO_STREAM, ALGO_TC, FLAG_KEEP_F32_MASTER, MEM_DEVICE, MEM_HOST, METRIC_IP, METRIC_L2, STORE_BF16,
 STORE_F16, STORE_F32) = (_cabi.ALGO_AUTO, _cabi.ALGO_SIMT, _cabi.ALGO_STREAM, _cabi.ALGO_TC, _cabi.FLAG_KEEP_F32_MASTER,
                          _cabi.MEM_DEVICE, _cabi.MEM_HOST, _cabi.METRIC_IP, _cabi.METRIC_L2, _cabi.STORE_BF16,
                          _cabi.STORE_F16, _cabi.STORE_F32)

_STORE_BY_NAME = {"f32": STORE_F32, "fp32": STORE_F32, "float32": STORE_F32,
                  "bf16": STORE_BF16, "bfloat16": STORE_BF16,
                  "f16": STORE_F16, "fp16": STORE_F16, "float16": STORE_F16}
_STORE_NAME = {STORE_F32: "f32", STORE_BF16: "bf16", STORE_F16: "f16"}
ALGO_BY_NAME = {"auto": ALGO_AUTO, "simt": ALGO_SIMT, "tc": ALGO_TC, "stream": ALGO_STREAM}


def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _locked(fn):
    """Selecting the stream (caller's torch stream vs the handle's own) and the C call that uses it must be one atomic
    step: two Python threads sharing an index (Flask's `predict`, app.py:351) could otherwise flip the stream between the
    other thread's selection and its launch.  ctypes releases the GIL inside the C call, so waiters do not spin."""
    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self._lock:
            return fn(self, *a, **kw)
    return wrapper


class _ReconstructCache:
    """The reference caller fetches every neighbour of a batch with its own ``index.reconstruct(i)`` call
    (``pipeline.py:491-509``: B x (K+10) calls, 3840 per training batch) -- on a GPU index each one is a device round
    trip.  After a host-path ``search`` the result ids are remembered; the FIRST ``reconstruct`` that follows gathers
    all of them with one kernel + one device-to-host copy (if that is at most ``_RC_MAX_BYTES``), and the calls are
    then served from host memory.  Rows are append-only, so cached rows never go stale; the cache is dropped at the
    next search.  Nothing is fetched unless ``reconstruct`` is actually called."""
    _RC_MAX_BYTES = 512 << 20

    def _rc_note_search(self, ids_host) -> None:
        self._rc_pending, self._rc_rows = ids_host, None

    def _rc_lookup(self, i: int):
        rows = getattr(self, "_rc_rows", None)
        if rows is None:
            pend = getattr(self, "_rc_pending", None)
            if pend is None:
                return None
            self._rc_pending = None
            ids = np.unique(pend[pend >= 0])
            if ids.size == 0 or ids.size * self.d * 4 > self._RC_MAX_BYTES:
                return None
            rows = self._rc_rows = (ids, self.reconstruct_batch(ids))
        ids, mat = rows
        j = int(np.searchsorted(ids, i))
        if j < ids.size and ids[j] == i:
            return mat[j].copy()
        return None


class FlatIndex(_ReconstructCache):
    """Exact flat nearest-neighbour index on one B200 (one row shard)."""

    is_trained = True          # faiss flat indexes need no training (vector_database.py:124)

    def __init__(self, d: int, metric: int = METRIC_L2, store="f32", device: Optional[int] = None,
                 keep_f32_master: bool = False, _handle=None):
        self._lib = _cabi.load()
        self._h = ctypes.c_void_p()
        self._lock = threading.RLock()
        self._stream_set = "own"
        self.nprobe = 1        # accepted and ignored: the flat index is exhaustive (vector_database.py:176-177)
        self._id_offset = 0
        if _handle is not None:
            self._h = _handle
        else:
            st = store if isinstance(store, int) else _STORE_BY_NAME[str(store).lower()]
            flags = FLAG_KEEP_F32_MASTER if keep_f32_master else 0
            _cabi.check(self._lib.rdb_create(int(d), int(metric), int(st), -1 if device is None else int(device),
                                             flags, ctypes.byref(self._h)))

    # ------------------------------------------------------------------ properties
    @property
    def ntotal(self) -> int:
        return int(self._lib.rdb_ntotal(self._h))

    @property
    def d(self) -> int:
        return int(self._lib.rdb_dim(self._h))

    @property
    def metric(self) -> int:
        return int(self._lib.rdb_metric(self._h))

    @property
    def metric_type(self) -> int:           # faiss numbering: METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1
        return 0 if self.metric == METRIC_IP else 1

    @property
    def store(self) -> str:
        return _STORE_NAME[int(self._lib.rdb_store_dtype(self._h))]

    @property
    def launch_count(self) -> int:
        return int(self._lib.rdb_launch_count(self._h))

    @property
    def host_sync_count(self) -> int:
        """Times a call on this index blocked the host on its stream (device-tensor searches must not)."""
        return int(self._lib.rdb_host_sync_count(self._h))

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        _cabi.check(rc, self._h)

    def _use_torch_stream(self, torch):
        s = torch.cuda.current_stream().cuda_stream
        if self._stream_set != s:
            self._check(self._lib.rdb_set_stream(self._h, ctypes.c_void_p(s)))
            self._stream_set = s

    def _use_own_stream(self):
        if self._stream_set != "own":
            self._check(self._lib.rdb_use_own_stream(self._h))
            self._stream_set = "own"

    def _as_host_f32(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"expected float32 [n, {self.d}], got {x.shape}")   # faiss asserts on d
        return x

    # ------------------------------------------------------------------ faiss surface
    def train(self, x) -> None:             # no-op (vector_database.py:128)
        return None

    def reserve(self, n_total: int) -> None:
        self._check(self._lib.rdb_reserve(self._h, int(n_total)))

    @_locked
    def add(self, x, normalize: bool = False) -> None:
        """index.add(x) (vector_database.py:138); ``normalize`` fuses _maybe_normalize (:100-105)."""
        if _is_cuda_tensor(x):
            import torch
            x = x.detach().to(torch.float32).contiguous()
            if x.dim() != 2 or x.shape[1] != self.d:
                raise RuntimeError(f"expected float32 [n, {self.d}], got {tuple(x.shape)}")
            self._use_torch_stream(torch)
            self._check(self._lib.rdb_add(self._h, ctypes.c_void_p(x.data_ptr()), x.shape[0], MEM_DEVICE,
                                          int(bool(normalize))))
            return
        x = self._as_host_f32(x)
        self._use_own_stream()
        self._check(self._lib.rdb_add(self._h, x.ctypes.data_as(ctypes.c_void_p), x.shape[0], MEM_HOST,
                                      int(bool(normalize))))

    @_locked
    def search(self, q, k: int, normalize: bool = False, algo=ALGO_AUTO, return_labels: bool = False):
        """index.search(q, k) -> (distances float32[nq,k], ids int64[nq,k]) best-first
        (vector_database.py:181).  torch CUDA queries give torch CUDA results."""
        k = int(k)
        algo = ALGO_BY_NAME[algo] if isinstance(algo, str) else int(algo)
        if _is_cuda_tensor(q):
            import torch
            q = q.detach().to(torch.float32).contiguous()
            if q.dim() != 2 or q.shape[1] != self.d:
                raise RuntimeError(f"expected float32 [nq, {self.d}], got {tuple(q.shape)}")
            nq = q.shape[0]
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            L = torch.empty((nq, k), dtype=torch.float32, device=q.device) if return_labels else None
            self._use_torch_stream(torch)
            self._check(self._lib.rdb_search_algo(
                self._h, ctypes.c_void_p(q.data_ptr()), nq, k, MEM_DEVICE, int(bool(normalize)), algo,
                ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                ctypes.c_void_p(L.data_ptr()) if L is not None else None))
            return (D, I, L) if return_labels else (D, I)
        q = self._as_host_f32(q)
        nq = q.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        L = np.empty((nq, k), dtype=np.float32) if return_labels else None
        self._use_own_stream()
        self._check(self._lib.rdb_search_algo(
            self._h, q.ctypes.data_as(ctypes.c_void_p), nq, k, MEM_HOST, int(bool(normalize)), algo,
            D.ctypes.data_as(ctypes.c_void_p), I.ctypes.data_as(ctypes.c_void_p),
            L.ctypes.data_as(ctypes.c_void_p) if L is not None else None))
        self._rc_note_search(I)
        return (D, I, L) if return_labels else (D, I)

    @_locked
    def reconstruct(self, i: int) -> np.ndarray:
        """index.reconstruct(i) -> float32[d] (pipeline.py:503)."""
        hit = self._rc_lookup(int(i))
        if hit is not None:
            return hit
        out = np.empty((self.d,), dtype=np.float32)
        self._use_own_stream()
        self._check(self._lib.rdb_reconstruct(self._h, int(i), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    @_locked
    def reconstruct_batch(self, ids):
        """Rows ``ids`` in one kernel; ids < 0 / out of range give zero rows (pipeline.py:511-512)."""
        if _is_cuda_tensor(ids):
            import torch
            ids = ids.detach().to(torch.int64).contiguous()
            out = torch.empty(tuple(ids.shape) + (self.d,), dtype=torch.float32, device=ids.device)
            self._use_torch_stream(torch)
            self._check(self._lib.rdb_reconstruct_batch(self._h, ctypes.c_void_p(ids.data_ptr()), ids.numel(),
                                                        MEM_DEVICE, ctypes.c_void_p(out.data_ptr())))
            return out
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.empty(ids.shape + (self.d,), dtype=np.float32)
        self._use_own_stream()
        self._check(self._lib.rdb_reconstruct_batch(self._h, ids.ctypes.data_as(ctypes.c_void_p), ids.size, MEM_HOST,
                                                    out.ctypes.data_as(ctypes.c_void_p)))
        return out

    # ------------------------------------------------------------------ additive API
    @_locked
    def set_labels(self, labels) -> None:
        lab = np.ascontiguousarray(labels, dtype=np.float32).reshape(-1)
        self._use_own_stream()
        self._check(self._lib.rdb_set_labels(self._h, lab.ctypes.data_as(ctypes.c_void_p), lab.size))

    def set_id_offset(self, offset: int) -> None:
        self._check(self._lib.rdb_set_id_offset(self._h, int(offset)))
        self._id_offset = int(offset)

    def set_option(self, name: str, value: int) -> None:
        """Per-handle tuning / test option (``rdb_set_option``): e.g. ``tc_cta_group``, ``tier1``, ``largek_scorer``."""
        self._check(self._lib.rdb_set_option(self._h, str(name).encode(), int(value)))

    def truncate(self, n_keep: int) -> None:
        """Forget the rows beyond the first ``n_keep`` (rolls a partially applied multi-shard add back)."""
        with self._lock:
            self._check(self._lib.rdb_truncate(self._h, int(n_keep)))
            self._rc_pending = self._rc_rows = None

    def release_scratch(self) -> None:
        """Free the grow-only search scratch; the index stays searchable (the scratch regrows on demand)."""
        with self._lock:
            self._check(self._lib.rdb_release_scratch(self._h))

    @_locked
    def search_shard(self, q, k: int, normalize: bool = False):
        """Per-shard candidates in merge form (torch CUDA in/out): (key, gid, labels, qnorm)."""
        import torch
        q = q.detach().to(torch.float32).contiguous()
        nq = q.shape[0]
        key = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        gid = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        lab = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        qn = torch.empty((nq,), dtype=torch.float32, device=q.device)
        self._use_torch_stream(torch)
        self._check(self._lib.rdb_search_shard(self._h, ctypes.c_void_p(q.data_ptr()), nq, int(k),
                                               int(bool(normalize)), ctypes.c_void_p(key.data_ptr()),
                                               ctypes.c_void_p(gid.data_ptr()), ctypes.c_void_p(lab.data_ptr()),
                                               ctypes.c_void_p(qn.data_ptr())))
        return key, gid, lab, qn

    @_locked
    def merge_shards(self, key, gid, lab, qnorm):
        """Merge [nq, nlists, k] candidate lists (torch CUDA) -> (D, I, L) as an unsharded search."""
        import torch
        nq, nlists, k = key.shape
        key, gid, lab = key.contiguous(), gid.contiguous(), lab.contiguous()
        D = torch.empty((nq, k), dtype=torch.float32, device=key.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=key.device)
        L = torch.empty((nq, k), dtype=torch.float32, device=key.device)
        self._use_torch_stream(torch)
        self._check(self._lib.rdb_merge_shards(self._h, ctypes.c_void_p(key.data_ptr()),
                                               ctypes.c_void_p(gid.data_ptr()), ctypes.c_void_p(lab.data_ptr()), nq,
                                               nlists, k, ctypes.c_void_p(qnorm.data_ptr()),
                                               ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                                               ctypes.c_void_p(L.data_ptr())))
        return D, I, L

    # ---- peer-memory exchange (CUDA IPC): buffers every rank of the node can map
    def ipc_alloc(self, nbytes: int):
        """Device buffer + its 64-byte CUDA IPC handle (bytes)."""
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        self._check(self._lib.rdb_ipc_alloc(self._h, int(nbytes), ctypes.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def ipc_open(self, handle: bytes) -> int:
        ptr = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.rdb_ipc_open(self._h, buf, ctypes.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, ptr: int) -> None:
        self._check(self._lib.rdb_ipc_close(self._h, ctypes.c_void_p(ptr)))

    def ipc_free(self, ptr: int) -> None:
        self._check(self._lib.rdb_ipc_free(self._h, ctypes.c_void_p(ptr)))

    @_locked
    def search_shard_into(self, q, k: int, normalize: bool, key_ptr: int, gid_ptr: int, lab_ptr: int, qnorm):
        """search_shard writing the candidates to raw device pointers (the IPC-exported buffer)."""
        import torch
        q = q.detach().to(torch.float32).contiguous()
        self._use_torch_stream(torch)
        self._check(self._lib.rdb_search_shard(self._h, ctypes.c_void_p(q.data_ptr()), q.shape[0], int(k),
                                               int(bool(normalize)), ctypes.c_void_p(key_ptr),
                                               ctypes.c_void_p(gid_ptr), ctypes.c_void_p(lab_ptr),
                                               ctypes.c_void_p(qnorm.data_ptr())))

    @_locked
    def merge_shards_peer(self, key_ptrs, gid_ptrs, lab_ptrs, nq: int, k: int, qnorm):
        """ONE kernel: gather list g from GPU g's memory over NVLink (P2P loads) while merging."""
        import torch
        G = len(key_ptrs)
        arr = lambda v: (ctypes.c_void_p * G)(*[ctypes.c_void_p(int(x)) for x in v])   # noqa: E731
        D = torch.empty((nq, k), dtype=torch.float32, device=qnorm.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=qnorm.device)
        L = torch.empty((nq, k), dtype=torch.float32, device=qnorm.device)
        self._use_torch_stream(torch)
        self._check(self._lib.rdb_merge_shards_peer(self._h, arr(key_ptrs), arr(gid_ptrs), arr(lab_ptrs), G, nq, k,
                                                    ctypes.c_void_p(qnorm.data_ptr()), ctypes.c_void_p(D.data_ptr()),
                                                    ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(L.data_ptr())))
        return D, I, L

    def enable_peer_access(self, peer_device: int) -> None:
        """Let this index's kernels read device memory of ``peer_device`` (single-process multi-GPU merge)."""
        self._check(self._lib.rdb_enable_peer_access(self._h, int(peer_device)))

    @_locked
    def copy_async(self, dst, src, on=None) -> None:
        """Stream-ordered copy ``dst <- src`` (contiguous CUDA tensors of equal byte size) by a kernel on THIS index's GPU
        and the calling thread's current stream of it (``rdb_copy_async``); either tensor may live on a peer GPU.  ``on`` =
        this index's torch device (default: ``dst.device``, i.e. the destination pulls; pass ``src.device`` for a push).
        A cross-device ``dst.copy_(src)`` is a DMA copy ordered against the source GPU's default stream instead."""
        import torch
        nbytes = dst.numel() * dst.element_size()
        if not (dst.is_contiguous() and src.is_contiguous()) or nbytes != src.numel() * src.element_size():
            raise RuntimeError("copy_async: contiguous tensors of equal byte size expected")
        with torch.cuda.device(dst.device if on is None else on):
            self._use_torch_stream(torch)
            self._check(self._lib.rdb_copy_async(self._h, ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                                 ctypes.c_size_t(nbytes)))

    @_locked
    def filter_first_k(self, idx, dist, lab, row_codes, excl_sorted, K: int, ntotal: Optional[int] = None):
        """Device-side rank-ordered exclusion + first-K compaction (pipeline.py:491-520); torch CUDA in/out.  ``ntotal`` =
        rows of the WHOLE database (row shards pass the global count; default: this index's own rows + id offset)."""
        import torch
        B, ks = idx.shape
        idx, dist, lab = idx.contiguous(), dist.contiguous(), lab.contiguous()
        oi = torch.empty((B, K), dtype=torch.int64, device=idx.device)
        od = torch.empty((B, K), dtype=torch.float32, device=idx.device)
        ol = torch.empty((B, K), dtype=torch.float32, device=idx.device)
        ne = 0 if excl_sorted is None else int(excl_sorted.numel())
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else None   # noqa: E731
        self._use_torch_stream(torch)
        if ntotal is None:
            ntotal = len(row_codes) if (ne and row_codes is not None) else self.ntotal + self._id_offset
        self._check(self._lib.rdb_filter_first_k(self._h, p(idx), p(dist), p(lab), B, ks, p(row_codes) if ne else None,
                                                 int(ntotal), p(excl_sorted) if ne else None, ne, int(K),
                                                 ctypes.c_void_p(oi.data_ptr()),
                                                 ctypes.c_void_p(od.data_ptr()), ctypes.c_void_p(ol.data_ptr())))
        return oi, od, ol

    @_locked
    def label_vote(self, labels_nq_k, kvote: int):
        if _is_cuda_tensor(labels_nq_k):
            import torch
            lab = labels_nq_k.contiguous()
            vote = torch.empty((lab.shape[0],), dtype=torch.float32, device=lab.device)
            self._use_torch_stream(torch)
            self._check(self._lib.rdb_label_vote(self._h, ctypes.c_void_p(lab.data_ptr()), lab.shape[0],
                                                 lab.shape[1], int(kvote), MEM_DEVICE,
                                                 ctypes.c_void_p(vote.data_ptr())))
            return vote
        lab = np.ascontiguousarray(labels_nq_k, dtype=np.float32)
        vote = np.empty((lab.shape[0],), dtype=np.float32)
        self._use_own_stream()
        self._check(self._lib.rdb_label_vote(self._h, lab.ctypes.data_as(ctypes.c_void_p), lab.shape[0],
                                             lab.shape[1], int(kvote), MEM_HOST,
                                             vote.ctypes.data_as(ctypes.c_void_p)))
        return vote

    def sync(self) -> None:
        self._check(self._lib.rdb_sync(self._h))

    def last_kernel_ms(self) -> Tuple[float, str, int]:
        ms, algo, ns = ctypes.c_float(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.rdb_last_kernel_ms(self._h, ctypes.byref(ms), ctypes.byref(algo), ctypes.byref(ns)))
        return float(ms.value), {ALGO_SIMT: "simt", ALGO_TC: "tc", ALGO_STREAM: "stream"}.get(algo.value, "?"), int(ns.value)

    @property
    def last_uncertified(self) -> int:
        """Queries of the last search that fell back from the certified tensor-core path to the exact kernel."""
        return int(self._lib.rdb_last_uncertified(self._h))

    @property
    def last_tier1(self) -> Tuple[int, int]:
        """(queries that entered the one-term certified pass, queries it could not certify) of the last search."""
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        self._check(self._lib.rdb_last_tier1(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return int(a.value), int(b.value)

    @property
    def last_tier1_candidates(self) -> int:
        """Candidates per query the one-term certified pass of the last search kept (0 = it did not run)."""
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        self._check(self._lib.rdb_last_tier1(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return int(c.value)

    def mem_info(self):
        """Device bytes of the stored rows / of the search scratch, and free / total of the device.  A closed index
        reports 0 / 0 (and the current device)."""
        a, s, b, c = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self._lib.rdb_mem_info(self._h if self._h else None, ctypes.byref(a), ctypes.byref(s), ctypes.byref(b),
                                           ctypes.byref(c)))
        return {"index_bytes": int(a.value), "scratch_bytes": int(s.value), "free": int(b.value), "total": int(c.value)}

    # ------------------------------------------------------------------ persistence (faiss IndexFlat layout)
    @_locked
    def save(self, path: str) -> None:
        self._use_own_stream()
        self._check(self._lib.rdb_serialize(self._h, str(path).encode()))

    @classmethod
    def load(cls, path: str, store="f32", device: Optional[int] = None, keep_f32_master: bool = False):
        lib = _cabi.load()
        h = ctypes.c_void_p()
        st = store if isinstance(store, int) else _STORE_BY_NAME[str(store).lower()]
        _cabi.check(lib.rdb_deserialize(str(path).encode(), int(st), -1 if device is None else int(device),
                                        FLAG_KEEP_F32_MASTER if keep_f32_master else 0, ctypes.byref(h)))
        return cls(0, _handle=h)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        """rdb_destroy: frees the rows, the scratch, the stream.  Idempotent."""
        with self._lock:
            h, self._h = self._h, ctypes.c_void_p()
            if h:
                self._lib.rdb_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
