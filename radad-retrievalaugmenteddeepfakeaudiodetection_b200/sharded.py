"""Row-sharded multi-GPU flat search: one process per GPU (``torch.distributed``), NCCL over NVLink.

The reference is single-GPU (``vector_database.py:23``); this is the B200-native scale-out of the same
``index.search``.  Top-k over a union of row sets equals the top-k of the per-set top-k's, so:

  1. database rows are split contiguously: rank g owns rows ``[g*ceil(N/G), (g+1)*ceil(N/G))``; global id =
     shard offset + local id, which preserves faiss insertion-order ids;
  2. queries are replicated; every rank runs the fused score+select kernel over its shard ->
     ``[nq, k]`` candidates in merge form (key, global id, label);
  3+4 fused (``exchange="peer"``, the default on NCCL): every rank leaves its candidates in a CUDA-IPC-mapped
     buffer; after a stream-ordered barrier (a 1-element all-reduce) ONE kernel on every rank merges the G lists
     while loading list g straight from GPU g's memory over NVLink/NVSwitch (P2P loads) -- no all-gather, no
     staging copy; nq*k*16 bytes per peer (10.5 MB at nq=65536, k=10);
  3, 4 separate (``exchange="nccl"``): one all-gather of the packed candidates, then the local merge kernel.  This
     path also runs on the ``gloo`` backend, so the exchange logic is covered by CPU tests.

Steps 2 and 4 are CUDA-only (no CPU fallback).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``: blocks of ceil(N/G) rows (last shards may be short or empty)."""
    per = -(-int(n_total) // int(world))
    start = min(int(n_total), rank * per)
    return start, min(int(n_total), start + per)


def pack_candidates(key, gid, lab):
    """(key f32[nq,k], gid i64[nq,k], lab f32[nq,k]) -> int32[nq,k,4] so one collective moves everything."""
    import torch
    return torch.cat([key.contiguous().view(torch.int32).unsqueeze(-1),
                      gid.contiguous().view(torch.int32).view(*gid.shape, 2),
                      lab.contiguous().view(torch.int32).unsqueeze(-1)], dim=-1).contiguous()


def unpack_candidates(packed):
    """int32[..., k, 4] -> (key f32[..., k], gid i64[..., k], lab f32[..., k])."""
    import torch
    key = packed[..., 0].contiguous().view(torch.float32)
    gid = packed[..., 1:3].contiguous().view(torch.int64).squeeze(-1)
    lab = packed[..., 3].contiguous().view(torch.float32)
    return key, gid, lab


def exchange_candidates(key, gid, lab, group=None):
    """All-gather the per-shard candidates.  Returns (key, gid, lab) shaped [nq, G, k] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    packed = pack_candidates(key, gid, lab)                      # [nq, k, 4]
    nq = packed.shape[0]
    out = torch.empty((world * nq,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)        # rank-major concatenation: [G * nq, k, 4]
    out = out.view(world, nq, *packed.shape[1:]).permute(1, 0, 2, 3).contiguous()   # [nq, G, k, 4]
    return unpack_candidates(out)


class ShardedFlatIndex:
    """Exact flat search over a database row-sharded across the ranks of a process group."""

    def __init__(self, d: int, metric: int, store="bf16", group=None, device: Optional[int] = None,
                 keep_f32_master: bool = False, exchange: str = "peer"):
        import torch.distributed as dist
        if __package__:
            from .flat_index import FlatIndex
        else:
            from flat_index import FlatIndex
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local = FlatIndex(d, metric, store, device=device, keep_f32_master=keep_f32_master)
        self.device_index = device
        self.ntotal_global = 0
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        self._peer = None          # (capacity_elems, own_ptr, [per-rank base ptrs]); two half-buffers, alternated
        self._peer_flip = 0
        self._barrier_token = None

    # ---- peer-memory exchange plumbing -------------------------------------------------------------------
    def _ensure_peer(self, nq: int, k: int, device):
        """(Re)allocate the IPC-exported candidate buffer: 2 halves x [nq*k] x (f32 key + i64 id + f32 label)."""
        import torch
        import torch.distributed as dist
        need = nq * k
        if self._peer is not None and self._peer[0] >= need:
            return
        if self._peer is not None:
            dist.barrier(self.group)
            for r, p in enumerate(self._peer[2]):
                if r != self.rank:
                    self.local.ipc_close(p)
            self.local.ipc_free(self._peer[1])
        cap = (max(need, 4096) + 3) & ~3       # the int64 id plane starts at base + 4 * cap: keep it 8-byte aligned (16 here)
        own, handle = self.local.ipc_alloc(2 * cap * 16)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        allh = torch.empty((self.world * 64,), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        allh = allh.cpu().view(self.world, 64)
        ptrs = [own if r == self.rank else self.local.ipc_open(bytes(allh[r].tolist())) for r in range(self.world)]
        self._peer = (cap, own, ptrs)
        self._barrier_token = torch.zeros((1,), dtype=torch.int32, device=device)

    @staticmethod
    def _planes(base: int, cap: int, half: int):
        b = base + half * cap * 16
        return b, b + cap * 4, b + cap * 12          # key f32 | gid i64 | label f32

    def set_shard(self, n_total_global: int) -> Tuple[int, int]:
        """Declare the global row count; returns this rank's [start, end) and fixes the id offset."""
        start, end = shard_bounds(n_total_global, self.world, self.rank)
        self.local.set_id_offset(start)
        self.local.reserve(max(end - start, 1))
        self.ntotal_global = int(n_total_global)
        return start, end

    def add_local(self, x, normalize: bool = False) -> None:
        """Append rows of THIS rank's shard (numpy or torch CUDA), in global row order."""
        self.local.add(x, normalize=normalize)

    def set_labels_local(self, labels) -> None:
        self.local.set_labels(labels)

    def search_from_host(self, q_host, k: int, normalize: bool = False):
        """Queries in (pinned) HOST memory, identical on every rank: each rank uploads only its 1/G slice over PCIe and
        the full batch is assembled with one all-gather over NVLink (G x less host->device traffic than G full
        uploads), then ``search``.  q_host: torch CPU float32 [nq, d].  Returns (D, I, L) on every rank (device)."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device_index if self.device_index is not None else torch.cuda.current_device())
        nq, d = q_host.shape
        if self.world == 1:
            return self.search(q_host.to(dev, non_blocking=True), k, normalize=normalize)
        per = -(-nq // self.world)
        lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        mine = torch.zeros((per, d), dtype=torch.float32, device=dev)
        if hi > lo:
            mine[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
        full = torch.empty((per * self.world, d), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return self.search(full[:nq], k, normalize=normalize)

    def search(self, q, k: int, normalize: bool = False):
        """q: torch CUDA [nq, d], identical on every rank.  Returns (D, I, L) on every rank."""
        if self.world > 1 and self.exchange == "peer":
            import torch
            import torch.distributed as dist
            nq = q.shape[0]
            self._ensure_peer(nq, k, q.device)
            cap, _, ptrs = self._peer
            half = self._peer_flip
            self._peer_flip ^= 1
            qn = torch.empty((nq,), dtype=torch.float32, device=q.device)
            kp, gp, lp = self._planes(ptrs[self.rank], cap, half)
            self.local.search_shard_into(q, k, normalize, kp, gp, lp, qn)
            # stream-ordered barrier: completes on this stream only after every rank's shard search has finished
            dist.all_reduce(self._barrier_token, group=self.group)
            planes = [self._planes(p, cap, half) for p in ptrs]
            # Every rank merges only ITS 1/G slice of the queries (G lists each, loaded straight from the peers' memory)
            # and the finished slices are all-gathered: G x less merge work and P2P traffic per rank than merging all nq
            # queries everywhere; the all-gather moves nq * k * 16 / G bytes per rank.
            per = -(-nq // self.world)
            lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
            D = torch.empty((per * self.world, k), dtype=torch.float32, device=q.device)
            I = torch.empty((per * self.world, k), dtype=torch.int64, device=q.device)
            L = torch.empty((per * self.world, k), dtype=torch.float32, device=q.device)
            mine = [t[self.rank * per:(self.rank + 1) * per] for t in (D, I, L)]
            if hi > lo:
                d, i, l = self.local.merge_shards_peer([p[0] + lo * k * 4 for p in planes], [p[1] + lo * k * 8 for p in planes],
                                                       [p[2] + lo * k * 4 for p in planes], hi - lo, k, qn[lo:hi])
                mine[0][:hi - lo].copy_(d)
                mine[1][:hi - lo].copy_(i)
                mine[2][:hi - lo].copy_(l)
            works = [dist.all_gather_into_tensor(full, part, group=self.group, async_op=True)
                     for full, part in zip((D, I, L), mine)]
            for w in works:
                w.wait()
            return D[:nq], I[:nq], L[:nq]
        key, gid, lab, qn = self.local.search_shard(q, k, normalize=normalize)
        if self.world == 1:
            return self.local.merge_shards(key.unsqueeze(1), gid.unsqueeze(1), lab.unsqueeze(1), qn)
        gkey, ggid, glab = exchange_candidates(key, gid, lab, self.group)
        return self.local.merge_shards(gkey, ggid, glab, qn)
