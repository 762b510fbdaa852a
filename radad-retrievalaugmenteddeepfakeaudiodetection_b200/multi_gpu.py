"""``MultiGpuFlatIndex`` -- one process, one host thread, G GPUs: the row-sharded index behind a single
``VectorDatabase``.

The reference constructs ONE ``VectorDatabase`` in ONE process (``pipeline.py:90``) and is single-GPU
(``vector_database.py:23``).  ``ShardedFlatIndex`` (sharded.py) scales the search out with one process per GPU under
``torchrun``; this class gives the same row-sharded search to an unmodified single-process caller: it has the faiss
flat-index duck type of ``FlatIndex`` (``add / search / reconstruct / ntotal / d / is_trained / train / nprobe``), so
``VectorDatabase`` and ``retrieve_similar_vectors`` work on top of it unchanged.

  add      every call is water-filled over the shards in contiguous pieces (global id = insertion order, exactly as
           faiss); each shard remembers its pieces as (local start, global start, count) segments.
  search   queries are copied to every GPU (P2P over NVLink), each GPU runs the fused score+select kernels on its own
           stream concurrently, local ids are mapped to global ids on the device, and ONE kernel on the primary GPU
           merges the G candidate lists while loading list g straight from GPU g's memory (P2P loads; the same fused
           gather+merge kernel the torchrun path uses) -- no all-gather, no host round trip.
  save     the faiss IndexFlat file layout with the rows in global id order (what ``FlatIndex.save`` writes), so a
           database written by G GPUs loads on one GPU and vice versa.

All arithmetic runs in ``libradad_flat.so``; torch is plumbing (device tensors, streams, events).  No CPU fallback.
"""
from __future__ import annotations

import bisect
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

if __package__:
    from ._cabi import ALGO_AUTO, METRIC_IP, METRIC_L2
    from .flat_index import FlatIndex, _ReconstructCache, _is_cuda_tensor
else:           # flat layout (see flat_index.py)
    from _cabi import ALGO_AUTO, METRIC_IP, METRIC_L2
    from flat_index import FlatIndex, _ReconstructCache, _is_cuda_tensor

_MIN_SPLIT_ROWS = 4096          # smaller add() calls go whole to the least-loaded shard (fewer segments)


class MultiGpuFlatIndex(_ReconstructCache):
    """Exact flat search over a database row-sharded across the GPUs of one box, driven by one process."""

    is_trained = True

    def __init__(self, d: int, metric: int = METRIC_L2, store="f32", devices: Optional[Sequence[int]] = None,
                 keep_f32_master: bool = False):
        import torch
        if devices is None or (isinstance(devices, str) and devices == "all"):
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(v) for v in devices]
        if not self.devices:
            raise RuntimeError("MultiGpuFlatIndex needs at least one CUDA device (no CPU fallback)")
        self._d, self._metric = int(d), int(metric)
        self.shards: List[FlatIndex] = [FlatIndex(d, metric, store, device=g, keep_f32_master=keep_f32_master)
                                        for g in self.devices]
        for a, ga in enumerate(self.devices):           # every GPU merges its slice of the queries from every shard's lists
            for gb in self.devices:
                if gb != ga:
                    self.shards[a].enable_peer_access(gb)
        self.nprobe = 1
        self._ntotal = 0
        self._pool = None                               # host threads driving the shards (fp32 stores only)
        self._cap = [0] * len(self.devices)            # rows reserved per shard (grown geometrically, before any add)
        # per shard: parallel lists of segment (local_start, global_start, count), ascending in both
        self._seg: List[List[Tuple[int, int, int]]] = [[] for _ in self.devices]
        self._seg_dev = [None] * len(self.devices)      # device copies (local_starts, deltas), rebuilt lazily
        self._glob = None                               # global routing table (starts, shard, delta), rebuilt lazily
        self.phase_probe = False                        # profiling aid: per-GPU host time of each search phase (syncs the streams)
        self.last_phases = None

    # ------------------------------------------------------------------ properties (FlatIndex duck type)
    @property
    def ntotal(self) -> int:
        return self._ntotal

    @property
    def d(self) -> int:
        return self._d

    @property
    def metric(self) -> int:
        return self._metric

    @property
    def metric_type(self) -> int:
        return 0 if self._metric == METRIC_IP else 1

    @property
    def store(self) -> str:
        return self.shards[0].store

    @property
    def launch_count(self) -> int:
        return sum(s.launch_count for s in self.shards)

    @property
    def last_uncertified(self) -> int:
        return sum(s.last_uncertified for s in self.shards)

    @property
    def shard_sizes(self) -> List[int]:
        return [s.ntotal for s in self.shards]

    def train(self, x) -> None:
        return None

    def reserve(self, n_total: int) -> None:
        per = -(-int(n_total) // len(self.shards))
        for g in range(len(self.shards)):
            self._reserve_shard(g, max(per, 1))

    def _reserve_shard(self, g: int, need: int) -> None:
        """Make room for `need` rows on shard g (x1.5 growth).  All shards of an add() are reserved BEFORE any row is
        stored, so an out-of-memory failure leaves the index exactly as it was (faiss `add` is all-or-nothing too)."""
        if need > self._cap[g]:
            cap = max(need, self._cap[g] + self._cap[g] // 2)
            self.shards[g].reserve(cap)
            self._cap[g] = cap

    # ------------------------------------------------------------------ ingest
    def _plan(self, n: int) -> List[Tuple[int, int]]:
        """Water-filling: how many rows of an n-row add() go to each shard, as (shard, count) in shard order."""
        sizes = self.shard_sizes
        if n < _MIN_SPLIT_ROWS:
            return [(int(np.argmin(sizes)), n)]
        target = -(-(sum(sizes) + n) // len(sizes))
        plan, left = [], n
        for g, sz in enumerate(sizes):
            take = min(left, max(0, target - sz))
            if take:
                plan.append((g, take))
                left -= take
        if left:                                        # rounding leftovers
            plan.append((int(np.argmin(sizes)), left))
        return plan

    def add(self, x, normalize: bool = False) -> None:
        """index.add(x) (vector_database.py:138): rows get global ids ntotal .. ntotal + n - 1."""
        import torch
        n = int(x.shape[0])
        if x.ndim != 2 or int(x.shape[1]) != self._d:
            raise RuntimeError(f"expected float32 [n, {self._d}], got {tuple(x.shape)}")
        plan = self._plan(n)
        want = {}
        for g, cnt in plan:
            want[g] = want.get(g, 0) + cnt
        for g, cnt in want.items():
            self._reserve_shard(g, self.shards[g].ntotal + cnt)
        # all-or-nothing (faiss `add` is; VectorDatabase.add_vectors_batch logs a failed slice and carries on with the
        # next one, vector_database.py:147-149): segments / ntotal are committed only after every piece is stored; a
        # failure truncates the shards that already took their piece, so no orphan rows and no reused global ids
        before = [s.ntotal for s in self.shards]
        staged, r0 = [], 0
        try:
            for g, cnt in plan:
                piece = x[r0:r0 + cnt]
                shard = self.shards[g]
                if _is_cuda_tensor(piece):
                    with torch.cuda.device(self.devices[g]):
                        shard.add(piece.to(torch.device("cuda", self.devices[g]), non_blocking=True), normalize=normalize)
                else:
                    shard.add(piece, normalize=normalize)
                staged.append((g, shard.ntotal - cnt, self._ntotal + r0, cnt))
                r0 += cnt
        except Exception:
            for g, shard in enumerate(self.shards):
                if shard.ntotal != before[g]:
                    shard.truncate(before[g])
            raise
        for g, local0, glob0, cnt in staged:
            seg = self._seg[g]
            if seg and seg[-1][0] + seg[-1][2] == local0 and seg[-1][1] + seg[-1][2] == glob0:
                seg[-1] = (seg[-1][0], seg[-1][1], seg[-1][2] + cnt)
            else:
                seg.append((local0, glob0, cnt))
            self._seg_dev[g] = None
        self._ntotal += n
        self._glob = None

    def set_labels(self, labels) -> None:
        lab = np.ascontiguousarray(labels, dtype=np.float32).reshape(-1)
        if lab.size != self._ntotal:
            raise RuntimeError(f"set_labels: {lab.size} labels for {self._ntotal} rows")
        for g, shard in enumerate(self.shards):
            if shard.ntotal:
                shard.set_labels(np.concatenate([lab[gs:gs + c] for (_, gs, c) in self._seg[g]]))

    # ------------------------------------------------------------------ id maps
    def _local_to_global(self, g: int, local_ids):
        """int64 local row ids (device tensor on shard g's GPU; -1 = empty slot) -> global ids."""
        import torch
        seg = self._seg[g]
        if len(seg) == 1:
            return torch.where(local_ids >= 0, local_ids + (seg[0][1] - seg[0][0]), local_ids)
        if self._seg_dev[g] is None:
            dev = local_ids.device
            self._seg_dev[g] = (torch.tensor([s[0] for s in seg], dtype=torch.int64, device=dev),
                                torch.tensor([s[1] - s[0] for s in seg], dtype=torch.int64, device=dev))
        starts, deltas = self._seg_dev[g]
        which = torch.bucketize(local_ids.clamp_min(0), starts, right=True) - 1
        return torch.where(local_ids >= 0, local_ids + deltas[which], local_ids)

    def _routing(self):
        if self._glob is None:
            rows = sorted((gs, g, ls - gs, c) for g, seg in enumerate(self._seg) for (ls, gs, c) in seg)
            self._glob = ([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows], {})
        return self._glob

    def _locate(self, gid: int) -> Tuple[int, int]:
        starts, shard, delta, cnt, _ = self._routing()
        j = bisect.bisect_right(starts, gid) - 1
        if gid < 0 or j < 0 or gid >= starts[j] + cnt[j]:
            raise RuntimeError(f"reconstruct: id {gid} out of range [0, {self._ntotal})")
        return shard[j], gid + delta[j]

    # ------------------------------------------------------------------ search
    def _workers(self):
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=len(self.shards), thread_name_prefix="rdb-shard")
        return self._pool

    def search(self, q, k: int, normalize: bool = False, algo=ALGO_AUTO, return_labels: bool = False):
        """index.search(q, k) -> (distances float32[nq,k], ids int64[nq,k]) best-first (vector_database.py:181);
        numpy in -> numpy out, torch CUDA in -> torch CUDA out (on the queries' device).

        Every GPU is driven by its own host thread (ctypes / torch release the GIL inside the driver calls):
          1. input   numpy: GPU g uploads only ITS 1/G slice of the batch over its own PCIe link, in parallel with the
                     others (a single upload to one GPU was 15 ms of a 99 ms step at C3 on 8 GPUs); CUDA tensor: one event
                     orders the workers behind the caller's stream;
          2. gather  the slices are exchanged GPU -> GPU over NVLink (P2P copies) so every GPU holds the whole batch;
          3. search  every GPU runs the fused score+select kernels over its row shard, local ids -> global ids;
          4. merge   GPU g merges ONLY the queries of its slice, loading the G candidate lists straight from the peers'
                     memory (the fused gather+merge kernel), and writes its slice of the result to the caller's buffers.
        """
        import threading
        import torch
        k = int(k)
        cuda_in = _is_cuda_tensor(q)
        if cuda_in:
            qsrc = q.detach().to(torch.float32).contiguous()
            if qsrc.dim() != 2 or qsrc.shape[1] != self._d:
                raise RuntimeError(f"expected float32 [nq, {self._d}], got {tuple(qsrc.shape)}")
        else:
            qsrc = np.ascontiguousarray(q, dtype=np.float32)
            if qsrc.ndim != 2 or qsrc.shape[1] != self._d:
                raise RuntimeError(f"expected float32 [nq, {self._d}], got {qsrc.shape}")
        nq = int(qsrc.shape[0])
        live = [g for g, s in enumerate(self.shards) if s.ntotal > 0]
        if not live:
            raise RuntimeError("search on an empty index")
        G = len(live)
        per = -(-nq // G) if nq else 0
        bounds = [(min(nq, t * per), min(nq, (t + 1) * per)) for t in range(G)]
        if cuda_in:
            outD = torch.empty((nq, k), dtype=torch.float32, device=qsrc.device)
            outI = torch.empty((nq, k), dtype=torch.int64, device=qsrc.device)
            outL = torch.empty((nq, k), dtype=torch.float32, device=qsrc.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(qsrc.device))          # the queries are final behind this event
        else:
            outD = np.empty((nq, k), dtype=np.float32)
            outI = np.empty((nq, k), dtype=np.int64)
            outL = np.empty((nq, k), dtype=np.float32)
            ready = None
        if nq == 0:
            return (outD, outI, outL) if return_labels else (outD, outI)
        peer_ok = cuda_in and qsrc.device.index in self.devices     # the shards' GPUs have each other's memory mapped
        barrier = threading.Barrier(G)
        slices, fulls, cands, events, errors = [None] * G, [None] * G, [None] * G, [None] * G, []
        probe = bool(self.phase_probe)
        import time as _time0
        t_start = _time0.perf_counter()
        phases = [dict() for _ in range(G)]

        def work(t):
            g = live[t]
            dev = torch.device("cuda", self.devices[g])
            lo, hi = bounds[t]
            import time as _time
            t_last = [_time.perf_counter()]

            def mark(name, st):
                # phase_probe = True / "sync": host time per phase with the stream drained at every phase boundary;
                # "events": no extra synchronisation -- host time stamps + CUDA events on the stream, resolved at the end
                if probe and self.phase_probe == "events":
                    ev_ = torch.cuda.Event(enable_timing=True)
                    ev_.record(st)
                    phases[t].setdefault("_ev", []).append((name, ev_, (_time.perf_counter() - t_start) * 1e3))
                elif probe:
                    st.synchronize()
                    now = _time.perf_counter()
                    phases[t][name] = (now - t_last[0]) * 1e3
                    t_last[0] = now
            try:
                with torch.cuda.device(dev):
                    st = torch.cuda.current_stream(dev)
                    mark("start", st)
                    # 1) this GPU's slice of the queries
                    if cuda_in:
                        st.wait_event(ready)
                        if qsrc.device == dev:
                            slices[t] = qsrc[lo:hi]
                        elif not peer_ok:
                            slices[t] = qsrc[lo:hi].to(dev, non_blocking=True)     # a GPU outside the index: no peer mapping
                        else:
                            # pulled by THIS worker on its own stream (a tensor .to() is enqueued on the source GPU's
                            # stream, behind whatever that GPU's worker has already launched)
                            slices[t] = torch.empty((hi - lo, self._d), dtype=torch.float32, device=dev)
                            if hi > lo:
                                self.shards[g].copy_async(slices[t], qsrc[lo:hi])
                    else:
                        slices[t] = torch.from_numpy(qsrc[lo:hi]).to(dev)          # pageable: returns when copied
                    ev = torch.cuda.Event()
                    ev.record(st)
                    events[t] = ev
                    mark("upload", st)
                    barrier.wait()
                    mark("barrier1", st)
                    # 2) the whole batch on this GPU: peers' slices over NVLink
                    if G == 1:
                        full = slices[t]
                    else:
                        full = torch.empty((nq, self._d), dtype=torch.float32, device=dev)
                        for u in range(G):
                            a, b = bounds[u]
                            if b > a:
                                st.wait_event(events[u])
                                self.shards[g].copy_async(full[a:b], slices[u])
                    fulls[t] = full
                    mark("gather", st)
                    # 3) search the shard
                    key, lid, lab, qn = self.shards[g].search_shard(full, k, normalize=normalize)
                    gid = self._local_to_global(g, lid)
                    ev2 = torch.cuda.Event()
                    ev2.record(st)
                    cands[t] = (key, gid, lab, qn, ev2)
                    mark("search", st)
                    barrier.wait()
                    mark("barrier2", st)
                    # 4) merge this GPU's slice of the queries from every shard's lists (P2P loads), hand it back
                    if hi > lo:
                        for u in range(G):
                            st.wait_event(cands[u][4])
                        esz = k
                        D, I, L = self.shards[g].merge_shards_peer(
                            [cands[u][0].data_ptr() + lo * esz * 4 for u in range(G)],
                            [cands[u][1].data_ptr() + lo * esz * 8 for u in range(G)],
                            [cands[u][2].data_ptr() + lo * esz * 4 for u in range(G)], hi - lo, k, qn[lo:hi])
                        if cuda_in and (outD.device == dev or not peer_ok):
                            outD[lo:hi].copy_(D, non_blocking=True)
                            outI[lo:hi].copy_(I, non_blocking=True)
                            outL[lo:hi].copy_(L, non_blocking=True)
                        elif cuda_in:
                            # pushed by this worker's GPU into the caller's tensors over NVLink (see copy_async)
                            self.shards[g].copy_async(outD[lo:hi], D, on=dev)
                            self.shards[g].copy_async(outI[lo:hi], I, on=dev)
                            self.shards[g].copy_async(outL[lo:hi], L, on=dev)
                        else:
                            torch.from_numpy(outD[lo:hi]).copy_(D)
                            torch.from_numpy(outI[lo:hi]).copy_(I)
                            torch.from_numpy(outL[lo:hi]).copy_(L)
                    st.synchronize()               # the peers' lists must stay alive until every merge has read them
                    mark("merge_download", st)
                    barrier.wait()
                    mark("barrier3", st)
            except BaseException as e:  # noqa: BLE001 - a failing worker must not leave the others at a barrier
                errors.append(e)
                barrier.abort()

        if G == 1:
            work(0)
        else:
            list(self._workers().map(work, range(G)))
        real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
        if real or errors:
            raise (real or errors)[0]
        if probe:
            for ph in phases:
                evs = ph.pop("_ev", None)
                if evs:
                    first = evs[0][1]
                    ph["host_ms_at_mark"] = {n: round(h, 3) for n, _, h in evs}
                    ph["gpu_ms_at_mark"] = {n: round(first.elapsed_time(e), 3) for n, e, _ in evs}
            self.last_phases = phases
        if not cuda_in:
            self._rc_note_search(outI)
        return (outD, outI, outL) if return_labels else (outD, outI)

    # ------------------------------------------------------------------ reconstruct
    def reconstruct(self, i: int) -> np.ndarray:
        hit = self._rc_lookup(int(i))
        if hit is not None:
            return hit
        g, local = self._locate(int(i))
        return self.shards[g].reconstruct(local)

    def reconstruct_batch(self, ids):
        """Rows ``ids`` (global); ids < 0 / out of range give zero rows (pipeline.py:511-512)."""
        import torch
        cuda_in = _is_cuda_tensor(ids)
        prim = torch.device("cuda", self.devices[0])
        t = ids.detach().to(torch.int64) if cuda_in else torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
        shape = tuple(t.shape)
        out_dev = t.device if cuda_in else None
        t = t.reshape(-1).to(prim)
        starts, shard, delta, cnt, cache = self._routing()
        if "dev" not in cache:
            mk = lambda v: torch.tensor(v, dtype=torch.int64, device=prim)     # noqa: E731
            cache["dev"] = (mk(starts), mk(shard), mk(delta), mk(cnt))
        st, sh, de, cn = cache["dev"]
        j = (torch.bucketize(t.clamp_min(0), st, right=True) - 1).clamp_min(0)
        valid = (t >= 0) & (t < st[j] + cn[j])
        which = torch.where(valid, sh[j], torch.full_like(t, -1))
        local = t + de[j]
        out = torch.zeros((t.numel(), self._d), dtype=torch.float32, device=prim)
        for g, s in enumerate(self.shards):
            if s.ntotal == 0:
                continue
            dev = torch.device("cuda", self.devices[g])
            lg = torch.where(which == g, local, torch.full_like(local, -1))
            with torch.cuda.device(dev):
                rows = s.reconstruct_batch(lg.to(dev, non_blocking=True))      # zero rows where lg == -1
            out += rows.to(prim, non_blocking=True)
        out = out.reshape(shape + (self._d,))
        if cuda_in:
            return out.to(out_dev)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ pass-throughs (elementwise on given tensors)
    def filter_first_k(self, idx, dist, lab, row_codes, excl_sorted, K: int, ntotal: Optional[int] = None):
        """Elementwise on the given tensors; ids are GLOBAL, so the bound is the global row count (a shard's own count
        would drop every neighbour that lives on another shard)."""
        import torch
        with torch.cuda.device(idx.device):
            own = [s for g, s in enumerate(self.shards) if self.devices[g] == idx.device.index]
            return (own[0] if own else self.shards[0]).filter_first_k(idx, dist, lab, row_codes, excl_sorted, K,
                                                                      ntotal=self._ntotal if ntotal is None else ntotal)

    def label_vote(self, labels_nq_k, kvote: int):
        return self.shards[0].label_vote(labels_nq_k, kvote)

    def sync(self) -> None:
        for s in self.shards:
            s.sync()

    def last_kernel_ms(self):
        r = [s.last_kernel_ms() for s in self.shards if s.ntotal]
        return max(v[0] for v in r), r[0][1], r[0][2]

    def set_option(self, name: str, value: int) -> None:
        for s in self.shards:
            s.set_option(name, value)

    def release_scratch(self) -> None:
        for s in self.shards:
            s.release_scratch()

    def mem_info(self):
        infos = [s.mem_info() for s in self.shards]
        return {"index_bytes": sum(i["index_bytes"] for i in infos), "scratch_bytes": sum(i["scratch_bytes"] for i in infos),
                "free": sum(i["free"] for i in infos),
                "total": sum(i["total"] for i in infos), "per_device": infos}

    # ------------------------------------------------------------------ persistence (faiss IndexFlat layout)
    def save(self, path: str) -> None:
        """fourcc IxF2 / IxFI, d, ntotal, 2 dummies, is_trained, metric, count, fp32 rows in GLOBAL id order."""
        cc = b"IxFI" if self._metric == METRIC_IP else b"IxF2"
        with open(path, "wb") as f:
            f.write(cc + struct.pack("<iqqqBiQ", self._d, self._ntotal, 1 << 20, 1 << 20, 1, self.metric_type,
                                     self._ntotal * self._d))
            chunk = max(1, (64 << 20) // (self._d * 4))
            for r0 in range(0, self._ntotal, chunk):
                ids = np.arange(r0, min(self._ntotal, r0 + chunk), dtype=np.int64)
                f.write(np.ascontiguousarray(self.reconstruct_batch(ids), dtype="<f4").tobytes())

    @classmethod
    def load(cls, path: str, store="f32", devices: Optional[Sequence[int]] = None, keep_f32_master: bool = False):
        with open(path, "rb") as f:
            head = f.read(37)
            if len(head) < 37 or head[:4] not in (b"IxF2", b"IxFI", b"IxFl"):
                raise RuntimeError(f"deserialize: {path} is not a faiss IndexFlat file")
            d, n, _, _, _, mt = struct.unpack("<iqqqBi", head[4:37])
            if mt > 1:
                f.read(4)                                   # metric_arg
            (count,) = struct.unpack("<Q", f.read(8))
            if d < 1 or n < 0 or count != n * d or mt not in (0, 1):
                raise RuntimeError(f"deserialize: {path} is not a faiss IndexFlat (L2 / IP) file")
            idx = cls(d, METRIC_IP if mt == 0 else METRIC_L2, store, devices, keep_f32_master)
            idx.reserve(n)
            per = -(-n // len(idx.shards)) if n else 0
            r0 = 0
            while r0 < n:                                   # shard-sized pieces: one segment per shard
                m = min(per, n - r0)
                step = max(1, (64 << 20) // (d * 4))
                g = int(np.argmin(idx.shard_sizes)) if idx._ntotal else 0
                for c0 in range(0, m, step):
                    c = min(step, m - c0)
                    rows = np.frombuffer(f.read(c * d * 4), dtype="<f4")
                    if rows.size != c * d:
                        raise RuntimeError(f"deserialize: {path} is truncated")
                    idx._add_to_shard(g, rows.reshape(c, d))
                r0 += m
        return idx

    def _add_to_shard(self, g: int, rows: np.ndarray) -> None:
        shard, cnt = self.shards[g], rows.shape[0]
        self._reserve_shard(g, shard.ntotal + cnt)
        shard.add(rows)
        local0 = shard.ntotal - cnt
        seg = self._seg[g]
        if seg and seg[-1][0] + seg[-1][2] == local0 and seg[-1][1] + seg[-1][2] == self._ntotal:
            seg[-1] = (seg[-1][0], seg[-1][1], seg[-1][2] + cnt)
        else:
            seg.append((local0, self._ntotal, cnt))
        self._seg_dev[g] = None
        self._ntotal += cnt
        self._glob = None

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_pool", None) is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for s in self.shards:
            s.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
