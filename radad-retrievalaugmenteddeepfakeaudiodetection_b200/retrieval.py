"""Device-resident ``retrieve_similar_vectors`` (reference ``pipeline.py:449-532``).

Same contract as the reference method -- rank-ordered self-exclusion by file basename, first K
survivors of the K+10 search results, zero / 0.0 / "" / NaN padding, same return arities -- but the
queries never leave the GPU: search, exclusion mask, compaction, the batched row gather (one kernel
instead of B*K ``index.reconstruct`` calls) and the label gather all run on the device, and the tensors
are handed straight to ``RADADModel``.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np


def _basename_codes(vector_db):
    """int64 code per database row identifying its file basename (cached on the VectorDatabase)."""
    paths = vector_db.vector_paths
    cache = getattr(vector_db, "_basename_cache", None)
    if cache is not None and cache[0] == len(paths) and cache[1] is paths:
        return cache[2], cache[3]
    table = {}
    codes = np.empty((len(paths),), dtype=np.int64)
    for i, p in enumerate(paths):
        b = os.path.basename(p)
        codes[i] = table.setdefault(b, len(table))
    vector_db._basename_cache = (len(paths), paths, codes, table)
    return codes, table


def retrieve_similar_vectors(vector_db, query_vectors, top_k: int, dim: Optional[int] = None,
                             query_paths: Optional[Sequence[str]] = None, exclude_self: bool = True,
                             return_info: bool = False, return_distances: bool = False,
                             training_file_ids: Optional[set] = None, device=None):
    """Returns ``(vec[B,K,D] f32, lbl[B,K] f32)`` + optionally ``paths: List[List[str]]`` and/or
    ``dist[B,K] f32`` -- all tensors on ``device`` (default: the queries' device)."""
    import torch

    q = query_vectors.detach()
    if not q.is_cuda:
        q = q.to(torch.device("cuda", vector_db.device_id))
    q = q.to(torch.float32)
    dev = q.device if device is None else torch.device(device)
    B, K = q.shape[0], int(top_k)
    D = int(dim if dim is not None else q.shape[1])

    def _pack(vec, lbl, pth, dst):
        if return_info and return_distances:
            return vec, lbl, pth, dst
        if return_info:
            return vec, lbl, pth
        if return_distances:
            return vec, lbl, dst
        return vec, lbl

    if vector_db.index is None or getattr(vector_db.index, "ntotal", 0) == 0:      # pipeline.py:465-476
        return _pack(torch.zeros(B, K, D, device=dev), torch.zeros(B, K, device=dev),
                     [[""] * K for _ in range(B)], torch.full((B, K), float("nan"), device=dev))

    # exclusion set -> sorted integer codes on the device (basenames are coded once per database, cached).  Prepared BEFORE
    # the search is launched: the host-to-device copy of the codes blocks the host on the stream, and behind the search it
    # would hold the filter / gather launches back until the search had finished
    row_codes = excl = None
    if exclude_self:
        codes_np, table = _basename_codes(vector_db)
        if query_paths is not None:
            names = {os.path.basename(p) for p in query_paths}
        else:
            names = set(training_file_ids or ())
        ex = np.fromiter((table[n] for n in names if n in table), dtype=np.int64)
        if ex.size:
            cache = getattr(vector_db, "_basename_codes_dev", None)
            if cache is None or cache[0] is not codes_np or cache[1].device != q.device:
                cache = (codes_np, torch.from_numpy(codes_np).to(q.device))
                vector_db._basename_codes_dev = cache
            row_codes = cache[1]
            excl = torch.from_numpy(np.sort(ex)).to(q.device)
    k_search = K + (10 if exclude_self else 0)                                      # pipeline.py:478
    try:
        dists, idxs, labs = vector_db.search_batch_with_labels(q, k=k_search)
    except Exception:  # noqa: BLE001 - pipeline.py:479-483: any search failure -> empty neighbours
        dists = torch.zeros(B, 0, device=q.device)
        idxs = torch.zeros(B, 0, dtype=torch.int64, device=q.device)
        labs = torch.zeros(B, 0, device=q.device)
    ks = idxs.shape[1]

    # first K survivors in rank order (pipeline.py:491-520): one CUDA kernel, one warp per query
    idx_k, dst_k, lbl_k = vector_db.index.filter_first_k(idxs, dists, labs, row_codes, excl, K)
    vec = vector_db.index.reconstruct_batch(idx_k)                                  # [B, K, D], zero rows for -1
    if vec.shape[2] != D:
        raise RuntimeError(f"index dimension {vec.shape[2]} != expected {D}")
    paths: List[List[str]] = []
    if return_info:
        host_idx = idx_k.cpu().numpy()
        vp = vector_db.vector_paths
        paths = [[vp[int(i)] if i >= 0 else "" for i in row] for row in host_idx]
    return _pack(vec.to(dev), lbl_k.to(dev, torch.float32), paths, dst_k.to(dev, torch.float32))
