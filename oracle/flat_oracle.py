"""CPU oracle for the RADAD retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package never does (it has
no CPU fallback and raises when the CUDA library is missing).

What this restates (all citations relative to /root/reference):

* ``vector_database.py`` in full -- the ``VectorDatabase`` wrapper (``:8-273``): lazy index
  creation, cosine normalisation, batched add with metadata bookkeeping, k clamping,
  search return types, save/load quirks.
* FAISS flat-index semantics (third-party dependency, NOT vendored in the reference:
  ``faiss-gpu-cu11==1.10.0`` / ``faiss-cpu`` unpinned, ``requirements.txt:7,11``).  The
  published algorithm of ``IndexFlatL2`` / ``IndexFlatIP``: exhaustive search, squared-L2
  ascending / inner product descending, 0-based int64 insertion-order ids, fp32 arithmetic.
  For nq >= 20 FAISS uses the BLAS expansion ``|q|^2 + |y|^2 - 2 q.y`` (negatives clamped
  to 0); below that a direct ``sum (q-y)^2`` loop.  Both are restated here.
* ``pipeline.py:449-532`` -- ``retrieve_similar_vectors`` (rank-ordered self-exclusion,
  padding, dtypes).

PARITY PINNING STATUS: **parity unpinned for the FAISS arithmetic** -- the reference ships
no tests / golden vectors / KATs and FAISS cannot be installed here.  What *is* pinned:
the wrapper / caller logic, by running the reference's own ``vector_database.py`` and
``pipeline.py`` source (imported from /root/reference with third-party modules stubbed)
on seeded inputs; see ``tests/golden/make_golden.py`` and the fixtures it wrote.

Tie rule of this oracle (FAISS leaves it implementation-defined): among equal distances
the LOWEST id wins and sorts first.  The CUDA kernels implement the same rule, so
integer-lattice inputs (exact arithmetic) must match bit-for-bit.
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

METRIC_L2 = 0
METRIC_IP = 1


# --------------------------------------------------------------------------------------
# numeric helpers
# --------------------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (round-to-nearest-even) and return as float32.

    Used for the bf16 configs: the oracle scores the SAME rounded values the kernels store
    (SURVEY 8d), so only summation order differs.  NaN is preserved; inf stays inf.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32).reshape(x.shape)
    nan = np.isnan(x)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out


def round_fp16(x: np.ndarray) -> np.ndarray:
    """Round float32 -> float16 -> float32 (what FAISS ``useFloat16`` storage does)."""
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def maybe_normalize(arr: np.ndarray, cosine: bool) -> np.ndarray:
    """vector_database.py:100-105 -- ``arr / (||arr||_2 + 1e-12)`` in float32, iff cosine."""
    if cosine:
        norms = np.linalg.norm(arr, axis=1, keepdims=True) + 1e-12
        arr = arr / norms
    return arr


def _select_topk_lowest_id(scores: np.ndarray, ids: np.ndarray, k: int,
                           largest: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Row-wise top-k of ``scores`` [nq, m] with ids [m] or [nq, m]; ties -> lowest id first.

    Returns (vals [nq,k], ids [nq,k]) sorted best-first.
    """
    nq, m = scores.shape
    if ids.ndim == 1:
        ids = np.broadcast_to(ids, (nq, m))
    key = -scores if largest else scores
    k = min(k, m)
    out_v = np.empty((nq, k), dtype=scores.dtype)
    out_i = np.empty((nq, k), dtype=np.int64)
    if m > 4 * k + 64:
        # partition first to keep the sort small, then take everything <= the k-th key so
        # that ties on the boundary are resolved by id and not by partition order.
        part = np.partition(key, k - 1, axis=1)[:, k - 1]
    else:
        part = None
    for r in range(nq):
        if part is not None:
            cand = np.nonzero(key[r] <= part[r])[0]
        else:
            cand = np.arange(m)
        order = np.lexsort((ids[r, cand], key[r, cand]))[:k]
        sel = cand[order]
        out_v[r] = scores[r, sel]
        out_i[r] = ids[r, sel]
    return out_v, out_i


# --------------------------------------------------------------------------------------
# FAISS flat index semantics
# --------------------------------------------------------------------------------------
class FlatIndexOracle:
    """Restatement of faiss.IndexFlatL2 / IndexFlatIP (the duck type used through
    ``VectorDatabase.index``: vector_database.py:138,181,169,210; pipeline.py:465,503).

    ``store`` models GPU storage precision: 'f32' (default), 'bf16' (this build's 16-bit
    store) or 'f16' (FAISS ``useFloat16``, vector_database.py:80).  With a 16-bit store the
    queries are rounded the same way before scoring (tensor-core operands), and
    ``reconstruct`` returns the rounded row up-converted to fp32.
    """

    def __init__(self, d: int, metric: int = METRIC_L2, store: str = "f32",
                 block_rows: int = 65536):
        self.d = int(d)
        self.metric = int(metric)
        self.store = store
        self.is_trained = True
        self._chunks: List[np.ndarray] = []
        self._xb: Optional[np.ndarray] = None
        self._block_rows = block_rows

    # ---- storage ----
    def _round(self, x: np.ndarray) -> np.ndarray:
        if self.store == "bf16":
            return round_bf16(x)
        if self.store == "f16":
            return round_fp16(x)
        return x

    @property
    def ntotal(self) -> int:
        n = sum(c.shape[0] for c in self._chunks)
        return n + (0 if self._xb is None else self._xb.shape[0])

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"add: expected [n,{self.d}] got {x.shape}")
        self._chunks.append(self._round(x).copy())

    def _base(self) -> np.ndarray:
        if self._chunks:
            parts = ([self._xb] if self._xb is not None else []) + self._chunks
            self._xb = np.concatenate(parts, axis=0) if len(parts) > 1 else parts[0]
            self._chunks = []
        if self._xb is None:
            return np.zeros((0, self.d), dtype=np.float32)
        return self._xb

    def reconstruct(self, i: int) -> np.ndarray:
        xb = self._base()
        i = int(i)
        if i < 0 or i >= xb.shape[0]:
            raise RuntimeError(f"reconstruct: id {i} out of range [0,{xb.shape[0]})")
        return xb[i].copy()

    def reconstruct_batch(self, ids: np.ndarray) -> np.ndarray:
        return self._base()[np.asarray(ids, dtype=np.int64)].copy()

    # ---- search ----
    def search(self, q: np.ndarray, k: int, direct: Optional[bool] = None
               ) -> Tuple[np.ndarray, np.ndarray]:
        """(distances float32[nq,k], ids int64[nq,k]) best-first.

        ``direct=None`` follows FAISS: BLAS expansion when nq >= 20, direct loops below.
        Slots beyond ntotal get id -1 and +inf (L2) / -inf (IP), as FAISS does.
        """
        q = self._round(np.ascontiguousarray(q, dtype=np.float32))
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"search: expected [nq,{self.d}] got {q.shape}")
        xb = self._base()
        nq, n = q.shape[0], xb.shape[0]
        k = int(k)
        largest = self.metric == METRIC_IP
        fill = -np.inf if largest else np.inf
        D = np.full((nq, k), fill, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        if n == 0 or k == 0 or nq == 0:
            return D, I
        if direct is None:
            direct = nq < 20
        kk = min(k, n)
        best_v = np.empty((nq, 0), dtype=np.float32)
        best_i = np.empty((nq, 0), dtype=np.int64)
        qn = np.einsum("ij,ij->i", q, q, dtype=np.float32) if not largest else None
        for s in range(0, n, self._block_rows):
            e = min(n, s + self._block_rows)
            yb = xb[s:e]
            if largest:
                sc = (q @ yb.T).astype(np.float32, copy=False)
            elif direct:
                # fvec_L2sqr: sum (q-y)^2 in fp32
                sc = np.empty((nq, e - s), dtype=np.float32)
                for r in range(nq):
                    diff = yb - q[r]
                    sc[r] = np.einsum("ij,ij->i", diff, diff, dtype=np.float32)
            else:
                yn = np.einsum("ij,ij->i", yb, yb, dtype=np.float32)
                ip = (q @ yb.T).astype(np.float32, copy=False)
                sc = (qn[:, None] + yn[None, :]) - np.float32(2.0) * ip
                np.maximum(sc, np.float32(0.0), out=sc)
            ids = np.arange(s, e, dtype=np.int64)
            v, i = _select_topk_lowest_id(sc, ids, kk, largest)
            cat_v = np.concatenate([best_v, v], axis=1)
            cat_i = np.concatenate([best_i, i], axis=1)
            best_v, best_i = _select_topk_lowest_id(cat_v, cat_i, kk, largest)
        D[:, :kk] = best_v
        I[:, :kk] = best_i
        return D, I

    def exact_scores(self, q: np.ndarray, ids: np.ndarray) -> np.ndarray:
        """float64 distances of q[r] to rows ids[r, :] (tie adjudication in the comparator)."""
        q = self._round(np.ascontiguousarray(q, dtype=np.float32)).astype(np.float64)
        xb = self._base()
        ids = np.asarray(ids, dtype=np.int64)
        out = np.empty(ids.shape, dtype=np.float64)
        for r in range(ids.shape[0]):
            rows = xb[np.clip(ids[r], 0, max(xb.shape[0] - 1, 0))].astype(np.float64)
            if self.metric == METRIC_IP:
                out[r] = rows @ q[r]
            else:
                diff = rows - q[r]
                out[r] = np.einsum("ij,ij->i", diff, diff)
        return out


# --------------------------------------------------------------------------------------
# the VectorDatabase wrapper, restated
# --------------------------------------------------------------------------------------
class OracleVectorDatabase:
    """Line-by-line CPU restatement of vector_database.py::VectorDatabase (:8-273) with
    ``FlatIndexOracle`` where the reference calls FAISS.  Same attributes, signatures,
    return types, error behaviour and quirks (``_cosine`` is not restored by ``load``)."""

    def __init__(self, config, store: str = "f32"):
        self.config = config                                             # :12
        self.index = None                                                # :13
        self.gpu_index = None
        self.vector_paths: list = []
        self.vector_labels: list = []
        self.vector_metadata: dict = {}
        self.db_path = os.path.join(config.vector_db_path, "faiss_index.bin")      # :18
        self.metadata_path = os.path.join(config.vector_db_path, "metadata.pkl")   # :19
        self.gpu_resources = None
        self.device_id = 0
        self._store = store
        os.makedirs(config.vector_db_path, exist_ok=True)                # :26

    def create_index(self, dimension: int):                              # :56-97
        index_type = self.config.vector_db_index_type.upper()
        if index_type == "L2":
            self.index = FlatIndexOracle(dimension, METRIC_L2, self._store)
        elif index_type == "IP":
            self.index = FlatIndexOracle(dimension, METRIC_IP, self._store)
        else:
            # IVF is out of scope (north star = exact flat); anything else: :72
            raise ValueError(f"Unsupported index type: {index_type}")
        self._cosine = (index_type == "IP") and bool(getattr(self.config, "normalize_for_ip", True))

    def _maybe_normalize(self, arr: np.ndarray) -> np.ndarray:           # :100-105
        return maybe_normalize(arr, getattr(self, "_cosine", False))

    def add_vectors_batch(self, vectors, paths, labels, metadata, batch_size: int = 10000):
        if vectors.shape[0] == 0:                                        # :110-112
            logging.warning("No vectors to add to database")
            return
        if self.index is None:                                           # :114-115
            self.create_index(vectors.shape[1])
        vectors = self._maybe_normalize(vectors.astype(np.float32, copy=False))
        vectors = np.ascontiguousarray(vectors)
        total = vectors.shape[0]
        added = 0
        for start in range(0, total, batch_size):                        # :134-149
            end = min(start + batch_size, total)
            batch = vectors[start:end]
            try:
                self.index.add(batch)
                added += batch.shape[0]
                self.vector_paths.extend(paths[start:end])
                self.vector_labels.extend(labels[start:end])
                for key, values in metadata.items():
                    self.vector_metadata.setdefault(key, [])
                    vals = values[start:end] if hasattr(values, "__getitem__") else [values] * len(batch)
                    self.vector_metadata[key].extend(vals)
            except Exception as e:  # noqa: BLE001 - mirrors the reference
                logging.error(f"Error adding batch {start}-{end}: {e}")
                continue
        logging.info(f"Added {added}/{total} vectors. Index ntotal={self.index.ntotal}")

    def add_vectors(self, vectors, paths, labels, metadata):             # :154-157
        batch_size = getattr(self.config, "vector_add_batch_size", 10000)
        self.add_vectors_batch(vectors, paths, labels, metadata, batch_size)

    def search_batch(self, query_vectors, k=None):                       # :159-182
        if self.index is None:
            raise ValueError("Vector database is empty. Build the database first.")
        k = int(k if k is not None else getattr(self.config, "top_k", 5))
        if query_vectors.ndim == 1:
            query_vectors = query_vectors.reshape(1, -1)
        query_vectors = self._maybe_normalize(query_vectors.astype(np.float32, copy=False))
        query_vectors = np.ascontiguousarray(query_vectors)
        k = min(k, self.index.ntotal)
        if k <= 0:
            logging.warning("No vectors available for search")
            return (np.zeros((len(query_vectors), 0), dtype=np.float32),
                    np.zeros((len(query_vectors), 0), dtype=np.int64))
        return self.index.search(query_vectors, k)

    def search(self, query_vector, k=None):                              # :185-188
        distances, indices = self.search_batch(query_vector.reshape(1, -1), k)
        return (distances[0] if len(distances) > 0 else np.array([]),
                indices[0] if len(indices) > 0 else np.array([]))

    def save(self):                                                      # :190-216
        try:
            if self.index is None:
                logging.warning("No index to save.")
                return
            np.save(self.db_path + ".npy", self.index._base())
            meta = {"paths": self.vector_paths, "labels": self.vector_labels,
                    "metadata": self.vector_metadata,
                    "index_type": self.config.vector_db_index_type,
                    "dimension": self.index.d}
            with open(self.metadata_path, "wb") as f:
                pickle.dump(meta, f)
        except Exception as e:  # noqa: BLE001
            logging.error(f"Error saving vector database: {e}")

    def load(self):                                                      # :218-242
        try:
            if not (os.path.exists(self.db_path + ".npy") and os.path.exists(self.metadata_path)):
                logging.warning("No saved vector database found")
                return
            with open(self.metadata_path, "rb") as f:
                meta = pickle.load(f)
            self.vector_paths = meta["paths"]
            self.vector_labels = meta["labels"]
            self.vector_metadata = meta["metadata"]
            xb = np.load(self.db_path + ".npy")
            metric = METRIC_IP if str(meta["index_type"]).upper() == "IP" else METRIC_L2
            idx = FlatIndexOracle(xb.shape[1], metric, self._store)
            idx.add(xb)
            self.index = idx            # NOTE: _cosine intentionally NOT restored (:218-242)
        except Exception as e:  # noqa: BLE001
            logging.error(f"Error loading vector database: {e}")


# --------------------------------------------------------------------------------------
# the caller, restated (pipeline.py:449-532) -- numpy in / numpy out
# --------------------------------------------------------------------------------------
def retrieve_similar_vectors_oracle(vector_db, q_np: np.ndarray, top_k: int, dim: int,
                                    query_paths: Optional[Sequence[str]] = None,
                                    exclude_self: bool = True,
                                    training_file_ids: Optional[set] = None):
    """Returns (vecs f32[B,K,D], labels f32[B,K], paths List[List[str]], dists f32[B,K])."""
    q_np = np.asarray(q_np, dtype=np.float32)
    B, K, D = q_np.shape[0], int(top_k), int(dim)
    exclude_ids = set()
    if exclude_self and query_paths is not None:                         # :461-463
        exclude_ids = {os.path.basename(p) for p in query_paths}
    if vector_db.index is None or getattr(vector_db.index, "ntotal", 0) == 0:   # :465-476
        return (np.zeros((B, K, D), np.float32), np.zeros((B, K), np.float32),
                [[""] * K for _ in range(B)], np.full((B, K), np.nan, np.float32))
    k_search = K + (10 if exclude_self else 0)                           # :478
    try:
        dists, idxs = vector_db.search_batch(q_np, k=k_search)
    except Exception:  # noqa: BLE001
        dists = np.zeros((B, 0), np.float32)
        idxs = np.zeros((B, 0), np.int64)
    vecs = np.zeros((B, K, D), np.float32)
    lbls = np.zeros((B, K), np.float32)
    dout = np.full((B, K), np.nan, np.float32)
    paths = [[""] * K for _ in range(B)]
    for r in range(B):                                                   # :491-520
        n = 0
        for ii, dd in zip(idxs[r], dists[r]):
            ii = int(ii)
            fname = os.path.basename(vector_db.vector_paths[ii])
            if exclude_self:
                if query_paths is not None:
                    if fname in exclude_ids:
                        continue
                elif fname in (training_file_ids or set()):
                    continue
            vecs[r, n] = vector_db.index.reconstruct(ii)
            lbls[r, n] = float(vector_db.vector_labels[ii])
            paths[r][n] = vector_db.vector_paths[ii]
            dout[r, n] = float(dd)
            n += 1
            if n == K:
                break
    return vecs, lbls, paths, dout


# --------------------------------------------------------------------------------------
# comparator (SURVEY 8d "Parity rule")
# --------------------------------------------------------------------------------------
def compare_topk(ours_d: np.ndarray, ours_i: np.ndarray, ref_d: np.ndarray, ref_i: np.ndarray,
                 exact_fn, metric: int, tol: float, abs_floor: float = 1e-6) -> Dict[str, float]:
    """Tolerance-aware comparison of two best-first result lists.

    ``exact_fn(ids[nq,m]) -> float64[nq,m]`` gives the oracle's exact distance of each
    query to arbitrary ids.  Rules:
      * every returned distance must match the exact distance of the returned id:
        ``|d_ours - d_exact| <= tol*|d_exact| + abs_floor``;
      * position p may differ from the reference only if the exact distance of our id is
        within tau of the reference distance at p (a tie group / boundary substitution);
      * no duplicates, ids in range; list must be sorted best-first within tau.
    Returns stats; raises AssertionError with a precise message on violation.
    """
    ours_d = np.asarray(ours_d)
    ours_i = np.asarray(ours_i)
    nq, k = ours_i.shape
    assert ref_i.shape[0] == nq and ref_i.shape[1] >= k, (ours_i.shape, ref_i.shape)
    exact_ours = exact_fn(ours_i)
    mism = 0
    inter = 0
    for r in range(nq):
        assert len(set(ours_i[r].tolist())) == k, f"query {r}: duplicate ids {ours_i[r]}"
        scale = max(abs(float(ref_d[r, k - 1])), 1e-30)
        tau = tol * scale + abs_floor
        de = exact_ours[r]
        err = np.abs(ours_d[r].astype(np.float64) - de)
        bad = err > tol * np.abs(de) + abs_floor
        assert not bad.any(), (f"query {r}: returned distance off: ours={ours_d[r][bad]} "
                               f"exact={de[bad]}")
        sgn = -1.0 if metric == METRIC_IP else 1.0
        srt = sgn * ours_d[r].astype(np.float64)
        assert (np.diff(srt) >= -tau).all(), f"query {r}: not sorted best-first: {ours_d[r]}"
        same = ours_i[r] == ref_i[r, :k]
        inter += len(set(ours_i[r].tolist()) & set(ref_i[r, :k].tolist()))
        if same.all():
            continue
        for p in np.nonzero(~same)[0]:
            mism += 1
            assert abs(de[p] - float(ref_d[r, p])) <= tau, (
                f"query {r} pos {p}: id {ours_i[r, p]} (exact {de[p]:.9g}) vs ref id "
                f"{ref_i[r, p]} (dist {float(ref_d[r, p]):.9g}); tau={tau:.3g}")
    return {"queries": nq, "k": k, "position_mismatches_within_tol": mism,
            "recall": inter / float(nq * k)}


# --------------------------------------------------------------------------------------
# CPU baseline used by bench.py (times the reference's algorithm on the host cores)
# --------------------------------------------------------------------------------------
def torch_cpu_flat_search(xb, q, k: int, metric: int, block_rows: int = 262144):
    """FAISS-style blocked sgemm + top-k on torch-CPU (all host threads).  Returns
    (D float32[nq,k], I int64[nq,k]).  Faster than numpy here; tie order is
    torch.topk's, so use ``FlatIndexOracle`` (not this) when adjudicating parity."""
    import torch

    xb_t = torch.as_tensor(xb)
    q_t = torch.as_tensor(q)
    largest = metric == METRIC_IP
    n = xb_t.shape[0]
    best_v = best_i = None
    qn = (q_t * q_t).sum(1, keepdim=True) if not largest else None
    for s in range(0, n, block_rows):
        yb = xb_t[s:s + block_rows]
        ip = q_t @ yb.T
        if largest:
            sc = ip
        else:
            sc = (qn + (yb * yb).sum(1)[None, :] - 2.0 * ip).clamp_(min=0.0)
        kk = min(k, sc.shape[1])
        v, i = torch.topk(sc, kk, dim=1, largest=largest, sorted=True)
        i = i + s
        if best_v is None:
            best_v, best_i = v, i
        else:
            cv = torch.cat([best_v, v], 1)
            ci = torch.cat([best_i, i], 1)
            kk = min(k, cv.shape[1])
            best_v, sel = torch.topk(cv, kk, dim=1, largest=largest, sorted=True)
            best_i = torch.gather(ci, 1, sel)
    return best_v.numpy(), best_i.numpy()
