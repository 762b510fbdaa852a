// Kernel 4 -- merge: fold L sorted candidate lists per query into the final best-first top-k, convert keys to
// the distances faiss returns (squared L2 ascending / inner product descending; vector_database.py:181),
// translate local row ids to global int64 ids, and gather the neighbour labels for the kNN label vote.
// The same kernel merges (a) the S per-chunk lists of one GPU and (b) the G per-shard lists after the
// all-gather of the multi-GPU path (IdxT = int64 there).
//
// Kernel 5 -- exact re-rank: re-score [Q, kc] candidates in fp32 against the fp32 master rows and sort them,
// which makes the split-precision tensor-core path return exact-fp32 neighbours.
#pragma once
#include "common.cuh"

namespace rdb {

constexpr int MERGE_LPL = 8;  // lists per lane -> up to 256 lists per query
constexpr int TC_PLAN_LISTS = 2;   // candidate lists per (query, chunk) of the tensor-core scorer (== TC_LISTS, score_tc.cuh)

struct MergeHead {
  uint32_t ok;     // ordered key, 0 = exhausted
  long long id;    // lower wins on equal key
};
__device__ __forceinline__ bool head_better(uint32_t ok_a, long long id_a, uint32_t ok_b, long long id_b) {
  return (ok_a > ok_b) || (ok_a == ok_b && id_a < id_b);
}

// key_in [Q][L][kc], idx_in [Q][L][kc] sorted best-first per list, idx < 0 = empty slot (only at list tails).
// lbl_in [Q][L][kc] (optional, labels that travelled with the candidates) or labels[] indexed by local id.
// metric_l2: dist = max(0, qnorm[q] - key) else dist = key.
// out_* [Q][kout]; missing results: id -1, dist +inf (L2) / -inf (IP), label 0 (faiss convention).
// One warp folds the L lists of query q.  Candidate loads are L2-coherent (ld.cg), so the function may also run in
// the kernel that produced the lists (the streaming scorer's last block) after a __threadfence().  kth_out (optional)
// receives the key of rank kout - 1 (-inf when fewer candidates exist); every out_* pointer may be null.
template <typename IdxT, bool STAGED = false, int LPL = MERGE_LPL>
__device__ __forceinline__ void merge_lists_warp(const float* key_in, const IdxT* idx_in, const float* lbl_in, int q, int L,
                                                 int kc, int kout, int metric_l2, const float* qnorm, long long id_offset,
                                                 const float* labels, float* out_dist, long long* out_idx,
                                                 float* out_lbl, float* out_key, int lane, float* kth_out,
                                                 int* sat_out = nullptr) {
  // sat_out (optional): 1 iff some list was consumed entirely (all kc entries valid and taken) -- the two-list cover of
  // the certified search (common.cuh, SelectSmall<KT, 2>) cannot vouch for rows such a list may have dropped
  // STAGED: key_in / idx_in point at this query's lists copied to shared memory (plain loads, base 0)
  const long long qbase = STAGED ? 0ll : (long long)q * L * kc;
  auto ld_idx = [&](long long o) -> long long { return STAGED ? (long long)idx_in[o] : (long long)__ldcg(idx_in + o); };
  auto ld_key = [&](long long o) -> float { return STAGED ? key_in[o] : __ldcg(key_in + o); };

  // Per list: the head and (prefetched up-front, so all loads of the prologue are independent) the entry behind it.
  // With many lists per query most lists contribute at most two results, so the dependent load after a win -- one L2
  // round trip per output rank -- is paid only from a list's third entry on.
  int ptr[LPL];
  uint32_t hok[LPL], sok[LPL];
  long long hid[LPL], sid[LPL];
  float hkey[LPL], skey[LPL];
#pragma unroll
  for (int i = 0; i < LPL; ++i) {
    ptr[i] = 0; hok[i] = 0; hid[i] = 0x7FFFFFFFFFFFFFFFll; hkey[i] = 0.f;
    sok[i] = 0; sid[i] = 0x7FFFFFFFFFFFFFFFll; skey[i] = 0.f;
    const int l = lane + 32 * i;
    if (l < L && kc > 0) {
      const long long o0 = qbase + (long long)l * kc;
      const long long id = ld_idx(o0);
      const float kv = ld_key(o0);
      long long id2 = -1; float kv2 = 0.f;
      if (kc > 1) { id2 = ld_idx(o0 + 1); kv2 = ld_key(o0 + 1); }
      if (id >= 0) { hkey[i] = kv; hok[i] = ordered_f32(kv); hid[i] = id; }
      if (id >= 0 && id2 >= 0) { skey[i] = kv2; sok[i] = ordered_f32(kv2); sid[i] = id2; }
    }
  }

  float kth = -CUDART_INF_F;
  for (int r = 0; r < kout; ++r) {
    // lane-local best head
    uint32_t bok = 0; long long bid = 0x7FFFFFFFFFFFFFFFll; int bi = 0;
#pragma unroll
    for (int i = 0; i < LPL; ++i)
      if (head_better(hok[i], hid[i], bok, bid)) { bok = hok[i]; bid = hid[i]; bi = i; }
    // warp arg-best
    uint32_t wok = bok; long long wid = bid; int wl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ook = __shfl_xor_sync(0xffffffffu, wok, o);
      const long long oid = __shfl_xor_sync(0xffffffffu, wid, o);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (head_better(ook, oid, wok, wid) || (ook == wok && oid == wid && ol < wl)) { wok = ook; wid = oid; wl = ol; }
    }
    const long long o = (long long)q * kout + r;
    if (wok == 0) {
      if (lane == 0) {
        if (out_dist) out_dist[o] = metric_l2 ? CUDART_INF_F : -CUDART_INF_F;
        if (out_idx) out_idx[o] = -1;
        if (out_lbl) out_lbl[o] = 0.f;
        if (out_key) out_key[o] = -CUDART_INF_F;
      }
      continue;
    }
    if (r == kout - 1) kth = unordered_f32(wok);
    if (lane == wl) {
      // emit + advance the winning list
      float kv = 0.f; int p = 0;
      uint32_t nok = 0; long long nid = 0x7FFFFFFFFFFFFFFFll; float nkey = 0.f;
#pragma unroll
      for (int i = 0; i < LPL; ++i) if (i == bi) { kv = hkey[i]; p = ptr[i]; nok = sok[i]; nid = sid[i]; nkey = skey[i]; }
      const int l = lane + 32 * bi;
      const long long src = qbase + (long long)l * kc + p;
      if (out_dist) {
        float d = kv;
        if (metric_l2) d = fmaxf(0.f, qnorm[q] - kv);
        out_dist[o] = d;
      }
      if (out_idx) out_idx[o] = wid + id_offset;   // wid is a local id when id_offset != 0, already global otherwise
      if (out_key) out_key[o] = kv;
      // labels[] indexed by row id: gathered AFTER the loop, all ranks at once (a random read per rank inside the
      // loop would stall every round for a DRAM round trip)
      if (out_lbl && (lbl_in || !labels || !out_idx)) out_lbl[o] = lbl_in ? __ldcg(lbl_in + (long long)q * L * kc + (src - qbase)) : 0.f;
      ++p;
      if (p >= 2) {                                // beyond the prefetched pair: dependent load
        nok = 0; nid = 0x7FFFFFFFFFFFFFFFll; nkey = 0.f;
        if (p < kc) {
          const long long id = ld_idx(src + 1);
          if (id >= 0) { nkey = ld_key(src + 1); nok = ordered_f32(nkey); nid = id; }
        }
      }
#pragma unroll
      for (int i = 0; i < LPL; ++i) if (i == bi) { ptr[i] = p; hok[i] = nok; hid[i] = nid; hkey[i] = nkey; }
    }
    __syncwarp();
  }
  if (out_lbl && out_idx && !lbl_in && labels) {
    __syncwarp();
    for (int r = lane; r < kout; r += 32) {
      const long long gid = __ldcg(out_idx + (long long)q * kout + r);
      out_lbl[(long long)q * kout + r] = gid >= 0 ? __ldg(labels + (gid - id_offset)) : 0.f;
    }
  }
  if (kth_out) *kth_out = kth;
  if (sat_out) {
    bool full = false;
#pragma unroll
    for (int i = 0; i < LPL; ++i) full = full || (kc > 0 && ptr[i] >= kc);
    const bool any = __any_sync(0xffffffffu, full);
    if (lane == 0) *sat_out = any ? 1 : 0;
  }
}

// ---- sorted 32-entry lists as one packed word per lane (lane j = rank j; key desc, id asc; 0 = empty slot) ----------
__device__ __forceinline__ unsigned long long pack_entry(float key, int idx) {
  return idx >= 0 ? pack_cand(key, uint32_t(idx)) : 0ull;
}
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
  return (static_cast<unsigned long long>(hi) << 32) | lo;
}
// best 32 of the union of two sorted lists, sorted: reverse one, lane-wise max (a bitonic sequence holding the best 32),
// five compare-exchange stages.  12 shuffles instead of up to 32 rounds of a warp arg-max.
__device__ __forceinline__ unsigned long long merge32(unsigned long long a, unsigned long long b, int lane) {
  const unsigned long long br = shfl64(b, 31 - lane);
  unsigned long long m = a > br ? a : br;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = shfl64(m, lane ^ o);
    const bool keep_small = (lane & o) != 0;
    m = ((m < x) == keep_small) ? m : x;
  }
  return m;
}
// full bitonic sort (descending) of one word per lane
__device__ __forceinline__ unsigned long long sort32(unsigned long long m, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int o = size >> 1; o > 0; o >>= 1) {
      const unsigned long long x = shfl64(m, lane ^ o);
      const bool desc = (lane & size) == 0 || size == 32;       // final pass: whole warp descending
      const bool keep_small = ((lane & o) != 0) == desc;
      m = ((m < x) == keep_small) ? m : x;
    }
  }
  return m;
}
// NW warp-held lists -> one (valid in warp 0 afterwards), log2(NW) levels through shared memory sx [NW][32]; the caller
// separates successive uses of sx with a block barrier.
template <int NW>
__device__ __forceinline__ unsigned long long block_tree_merge32(unsigned long long m, unsigned long long* sx, int warp,
                                                                 int lane) {
#pragma unroll
  for (int s = 1; s < NW; s <<= 1) {
    if ((warp & (2 * s - 1)) == s) sx[warp * 32 + lane] = m;
    __syncthreads();
    if ((warp & (2 * s - 1)) == 0) m = merge32(m, sx[(warp + s) * 32 + lane], lane);
  }
  return m;
}

// Few queries, many short lists (k <= 32; the mid-batch regime -- 256 queries against 1M rows leave 74 lists of 15 per
// query): one BLOCK per query.  Every warp folds its share of the lists with merge32 (all loads issued up-front), the
// warps' partial lists fold through shared memory, warp 0 emits.  One warp per query (merge_lists_kernel) walks k rounds
// of dependent arg-best steps over all L heads: 19.7 us for that case, with 256 warps on the whole GPU.
constexpr int MERGE_TREE_WARPS = 8;
static __global__ void __launch_bounds__(MERGE_TREE_WARPS * 32) merge_lists_tree_kernel(
    const float* __restrict__ key_in, const int* __restrict__ idx_in, int Q, int L, int kc, int kout, int metric_l2,
    const float* __restrict__ qnorm, long long id_offset, const float* __restrict__ labels, float* __restrict__ out_dist,
    long long* __restrict__ out_idx, float* __restrict__ out_lbl, float* __restrict__ out_key,
    const int* __restrict__ run_if, int* __restrict__ sat = nullptr) {
  __shared__ unsigned long long sx[MERGE_TREE_WARPS * 32];
  __shared__ unsigned long long s_last[MERGE_TREE_WARPS];   // per warp: best LAST entry of its full lists (sat)
  if (run_if && __ldcg(run_if) == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    unsigned long long acc = 0ull, best_last = 0ull;
    for (int l0 = warp; l0 < L; l0 += MERGE_TREE_WARPS * 4) {
      float kk[4]; int ii[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int li = l0 + MERGE_TREE_WARPS * j;
        const bool ok = li < L && lane < kc;
        const long long o = ((long long)q * L + (ok ? li : 0)) * kc + (ok ? lane : 0);
        kk[j] = __ldcg(key_in + o); ii[j] = ok ? __ldcg(idx_in + o) : -1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned long long e = pack_entry(kk[j], ii[j]);
        // merge32 needs (key desc, id asc) order; a producer that emits equal keys in another order gets its list sorted
        const unsigned long long nx = shfl64(e, min(lane + 1, 31));
        if (__any_sync(0xffffffffu, e < nx)) e = sort32(e, lane);
        const unsigned long long last = shfl64(e, kc - 1);          // non-zero iff the list is full
        best_last = last > best_last ? last : best_last;
        acc = merge32(acc, e, lane);
      }
    }
    if (lane == 0) s_last[warp] = best_last;
    const unsigned long long m = block_tree_merge32<MERGE_TREE_WARPS>(acc, sx, warp, lane);
    if (sat && warp == 0) {
      // a full list whose last entry made it into the result was consumed entirely (entries are distinct words)
      const unsigned long long thr = shfl64(m, kout - 1);
      const unsigned long long bl = lane < MERGE_TREE_WARPS ? s_last[lane] : 0ull;
      const bool any = __any_sync(0xffffffffu, bl != 0ull && bl >= thr);
      if (lane == 0) sat[q] = any ? 1 : 0;
    }
    if (warp == 0 && lane < kout) {
      const float kv = m ? unordered_f32(uint32_t(m >> 32)) : -CUDART_INF_F;
      const int mi = m ? int(0xFFFFFFFFu - uint32_t(m)) : -1;
      const long long o = (long long)q * kout + lane;
      if (out_dist) out_dist[o] = mi < 0 ? (metric_l2 ? CUDART_INF_F : -CUDART_INF_F)
                                         : (metric_l2 ? fmaxf(0.f, qnorm[q] - kv) : kv);
      if (out_idx) out_idx[o] = mi < 0 ? -1ll : (long long)mi + id_offset;
      if (out_key) out_key[o] = kv;
      if (out_lbl) out_lbl[o] = (mi >= 0 && labels) ? __ldg(labels + mi) : 0.f;
    }
    __syncthreads();
  }
}

template <typename IdxT, int LPL = MERGE_LPL>
__global__ void __launch_bounds__(128) merge_lists_kernel(const float* __restrict__ key_in,
                                                          const IdxT* __restrict__ idx_in,
                                                          const float* __restrict__ lbl_in, int Q, int L, int kc,
                                                          int kout, int metric_l2, const float* __restrict__ qnorm,
                                                          long long id_offset, const float* __restrict__ labels,
                                                          float* __restrict__ out_dist,
                                                          long long* __restrict__ out_idx,
                                                          float* __restrict__ out_lbl,
                                                          float* __restrict__ out_key,
                                                          const int* __restrict__ run_if = nullptr,
                                                          int stage_bytes_per_warp = 0,
                                                          const int* __restrict__ q_dev = nullptr,
                                                          const int* __restrict__ l_dev = nullptr,
                                                          int* __restrict__ sat = nullptr) {
  const int lane = threadIdx.x & 31;
  if (q_dev) Q = min(Q, __ldcg(q_dev));        // device-sized launch (DevPlan): a small fixed grid walks the queries
  if (l_dev) L = __ldcg(l_dev);
  if (run_if && __ldcg(run_if) == 0) return;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Q; q += nwarps) {
    if (stage_bytes_per_warp > 0) {
      // long merges (many lists x large k): every output rank would otherwise pay a dependent L2 round trip after a
      // list's second entry.  Copy this query's L x kc candidates to shared memory once (coalesced) and merge from there.
      extern __shared__ __align__(16) unsigned char merge_smem[];
      unsigned char* mine = merge_smem + size_t(threadIdx.x >> 5) * stage_bytes_per_warp;
      const int n = L * kc;
      IdxT* sidx = reinterpret_cast<IdxT*>(mine);                                     // [n] ids first (8-byte aligned)
      float* skey = reinterpret_cast<float*>(mine + size_t(n) * sizeof(IdxT));        // [n] keys
      const long long qbase = (long long)q * n;
      __syncwarp();
      for (int i = lane; i < n; i += 32) { sidx[i] = __ldcg(idx_in + qbase + i); skey[i] = __ldcg(key_in + qbase + i); }
      __syncwarp();
      merge_lists_warp<IdxT, true, LPL>(skey, sidx, lbl_in, q, L, kc, kout, metric_l2, qnorm, id_offset, labels, out_dist,
                                        out_idx, out_lbl, out_key, lane, nullptr, sat ? sat + q : nullptr);
    } else {
      merge_lists_warp<IdxT, false, LPL>(key_in, idx_in, lbl_in, q, L, kc, kout, metric_l2, qnorm, id_offset, labels,
                                         out_dist, out_idx, out_lbl, out_key, lane, nullptr, sat ? sat + q : nullptr);
    }
  }
}

// A single, already sorted list per query (the k > 128 path with one row chunk): nothing to merge -- convert keys to
// distances, local to global ids and gather the labels, one thread per output slot.  Same output conventions as
// merge_lists_warp (missing results: id -1, +inf / -inf, label 0, key -inf).
static __global__ void __launch_bounds__(256) finalize_sorted_list_kernel(const float* __restrict__ key_in,
                                                                   const int* __restrict__ idx_in, int Q, int k,
                                                                   int metric_l2, const float* __restrict__ qnorm,
                                                                   long long id_offset, const float* __restrict__ labels,
                                                                   float* __restrict__ out_dist,
                                                                   long long* __restrict__ out_idx,
                                                                   float* __restrict__ out_lbl,
                                                                   float* __restrict__ out_key) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Q * k) return;
  const int q = int(t / k);
  const int id = idx_in[t];
  const float kv = key_in[t];
  if (id < 0) {
    if (out_dist) out_dist[t] = metric_l2 ? CUDART_INF_F : -CUDART_INF_F;
    if (out_idx) out_idx[t] = -1;
    if (out_lbl) out_lbl[t] = 0.f;
    if (out_key) out_key[t] = -CUDART_INF_F;
    return;
  }
  if (out_dist) out_dist[t] = metric_l2 ? fmaxf(0.f, qnorm[q] - kv) : kv;
  if (out_idx) out_idx[t] = (long long)id + id_offset;
  if (out_key) out_key[t] = kv;
  if (out_lbl) out_lbl[t] = labels ? __ldg(labels + id) : 0.f;
}

// host helper: bytes of dynamic shared memory per warp for the staged form, or 0 when staging does not pay / fit
template <typename IdxT>
inline size_t merge_stage_bytes(int L, int kc, int kout) {
  const size_t b = (size_t(L) * kc * (sizeof(IdxT) + 4) + 15) & ~size_t(15);
  // (kout >= 8: the reference's own k_search = 15 with 148 lists -- Q = 256 against 1M rows -- took 20 us unstaged)
  return (kout >= 8 && L >= 4 && b <= 24 * 1024) ? b : 0;
}

// Exact fp32 re-rank.  One warp per query.
//   cand_idx [Q][kc] local ids from the approximate pass (best-first, -1 = empty), cand_key [Q][kc] approximate keys
//   qf [Q, D] fp32 (already normalised), master [N, D] fp32, ynorm [N]
// Writes exact keys in best-first order (key desc, id asc) into out_key/out_idx [Q][kc] and a per-query
// certificate.  Let B = eps * |q| * max|y| (x2 for the L2 key 2 q.y - |y|^2) bound |approx - exact|.  A row outside
// the candidate set has approx key <= the worst candidate's approx key, hence exact key <= approx_worst + B.  If the
// exact kout-th key is strictly above that, no outside row can enter the exact top-kout:
//   cert[q] = 1  iff  exact_key[kout-1] > approx_worst + B        (or the candidate set holds every row).
// Uncertified queries are appended to `uncert_list` (count in uncert_count) for the exact fallback.
constexpr int RERANK_MAX_KC = 128;
constexpr int RERANK_THREADS = 256;   // 8 warps share a query's candidates (4 rows each at kc = 32)
// One BLOCK (8 warps) per query: the warps share the candidates (each exact dot product runs 4 / 8 independent 128-bit
// loads per lane deep, so long rows -- the reference's D = 5376 -- are bandwidth- not latency-bound), then the block
// ranks the kc exact keys by counting and thread 0 evaluates the certificate.
template <bool L2>
__global__ void __launch_bounds__(RERANK_THREADS) rerank_exact_kernel(const long long* __restrict__ cand_idx,
                                                                      const float* __restrict__ cand_key, int Q, int kc,
                                                                      int kout, const float* __restrict__ qf,
                                                                      const float* __restrict__ master,
                                                                      const float* __restrict__ ynorm, int D, float eps,
                                                                      const float* __restrict__ qnorm,
                                                                      const float* __restrict__ ynorm_max,
                                                                      long long ntotal, float* __restrict__ out_key,
                                                                      long long* __restrict__ out_idx,
                                                                      int* __restrict__ uncert_list,
                                                                      int* __restrict__ uncert_count,
                                                                      const float* __restrict__ qres = nullptr,
                                                                      const float* __restrict__ yres_max = nullptr,
                                                                      float accum_eps = 0.f,
                                                                      const int* __restrict__ q_dev = nullptr,
                                                                      const int* __restrict__ qmap = nullptr,
                                                                      // final form of the first kout ranks, written by
                                                                      // this kernel (was a separate one-list merge launch):
                                                                      // distances / global ids / labels / raw keys
                                                                      float* __restrict__ fin_dist = nullptr,
                                                                      long long* __restrict__ fin_idx = nullptr,
                                                                      float* __restrict__ fin_lbl = nullptr,
                                                                      float* __restrict__ fin_key = nullptr,
                                                                      long long id_offset = 0,
                                                                      const float* __restrict__ labels = nullptr,
                                                                      // two-list cover: 1 = a candidate list was
                                                                      // consumed entirely -> never certified
                                                                      const int* __restrict__ sat = nullptr) {
  __shared__ uint32_t s_ok[RERANK_MAX_KC];     // ordered exact keys (0 = empty slot)
  __shared__ long long s_id[RERANK_MAX_KC];
  __shared__ float s_kth;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (q_dev) Q = min(Q, __ldcg(q_dev));        // device-sized launch: the grid covers the worst case
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
  const float* qr = qf + (long long)q * D;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(qr) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(master) & 15) == 0);
  for (int j = w; j < kc; j += RERANK_THREADS / 32) {
    const long long id = cand_idx[(long long)q * kc + j];
    if (id < 0) {
      if (lane == 0) { s_ok[j] = 0u; s_id[j] = 0x7FFFFFFFFFFFFFFFll; }
      continue;
    }
    const float* yr = master + id * (long long)D;
    float s = 0.f;
    if (vec4) {
      const float4* q4 = reinterpret_cast<const float4*>(qr);
      const float4* y4 = reinterpret_cast<const float4*>(yr);
      const int n4 = D >> 2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int c = lane;
      // long rows (the reference's D = 5376): eight independent 128-bit loads per lane and round trip; the four
      // accumulators receive their elements in the same order as in the 4-deep loop below, so keys are bit-identical
      for (; c + 224 < n4; c += 256) {
        float4 y[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) y[u] = __ldg(y4 + c + 32 * u);
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = q4[c + 32 * u];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const float4 y0 = y[4 * hlf], y1 = y[4 * hlf + 1], y2 = y[4 * hlf + 2], y3 = y[4 * hlf + 3];
          const float4 x0 = x[4 * hlf], x1 = x[4 * hlf + 1], x2 = x[4 * hlf + 2], x3 = x[4 * hlf + 3];
          a0 = fmaf(x0.x, y0.x, a0); a0 = fmaf(x0.y, y0.y, a0); a0 = fmaf(x0.z, y0.z, a0); a0 = fmaf(x0.w, y0.w, a0);
          a1 = fmaf(x1.x, y1.x, a1); a1 = fmaf(x1.y, y1.y, a1); a1 = fmaf(x1.z, y1.z, a1); a1 = fmaf(x1.w, y1.w, a1);
          a2 = fmaf(x2.x, y2.x, a2); a2 = fmaf(x2.y, y2.y, a2); a2 = fmaf(x2.z, y2.z, a2); a2 = fmaf(x2.w, y2.w, a2);
          a3 = fmaf(x3.x, y3.x, a3); a3 = fmaf(x3.y, y3.y, a3); a3 = fmaf(x3.z, y3.z, a3); a3 = fmaf(x3.w, y3.w, a3);
        }
      }
      for (; c + 96 < n4; c += 128) {
        const float4 y0 = __ldg(y4 + c), y1 = __ldg(y4 + c + 32), y2 = __ldg(y4 + c + 64), y3 = __ldg(y4 + c + 96);
        const float4 x0 = q4[c], x1 = q4[c + 32], x2 = q4[c + 64], x3 = q4[c + 96];
        a0 = fmaf(x0.x, y0.x, a0); a0 = fmaf(x0.y, y0.y, a0); a0 = fmaf(x0.z, y0.z, a0); a0 = fmaf(x0.w, y0.w, a0);
        a1 = fmaf(x1.x, y1.x, a1); a1 = fmaf(x1.y, y1.y, a1); a1 = fmaf(x1.z, y1.z, a1); a1 = fmaf(x1.w, y1.w, a1);
        a2 = fmaf(x2.x, y2.x, a2); a2 = fmaf(x2.y, y2.y, a2); a2 = fmaf(x2.z, y2.z, a2); a2 = fmaf(x2.w, y2.w, a2);
        a3 = fmaf(x3.x, y3.x, a3); a3 = fmaf(x3.y, y3.y, a3); a3 = fmaf(x3.z, y3.z, a3); a3 = fmaf(x3.w, y3.w, a3);
      }
      for (; c < n4; c += 32) {
        const float4 y0 = __ldg(y4 + c);
        const float4 x0 = q4[c];
        a0 = fmaf(x0.x, y0.x, a0); a0 = fmaf(x0.y, y0.y, a0); a0 = fmaf(x0.z, y0.z, a0); a0 = fmaf(x0.w, y0.w, a0);
      }
      s = (a0 + a1) + (a2 + a3);
    } else {
      for (int c = lane; c < D; c += 32) s = fmaf(qr[c], __ldg(yr + c), s);
    }
    s = warp_sum(s);
    const float key = L2 ? fmaf(2.0f, s, -ynorm[id]) : s;
    if (lane == 0) { s_ok[j] = ordered_f32(key); s_id[j] = id; }
  }
  if (threadIdx.x == 0) s_kth = -CUDART_INF_F;
  __syncthreads();
  // rank by counting (kc <= 128 = one candidate per thread): valid candidates by (key desc, id asc), empty slots
  // after them in slot order
  const int i = threadIdx.x;
  const bool in = i < kc;
  const uint32_t mok = in ? s_ok[i] : 0u;
  const long long mid = in ? s_id[i] : 0x7FFFFFFFFFFFFFFFll;
  const bool empty = in && mok == 0u;
  int nvalid = 0, empties_before = 0;
  for (int j = 0; j < kc; ++j) {
    const bool ej = s_ok[j] == 0u;
    nvalid += ej ? 0 : 1;
    empties_before += (ej && j < i) ? 1 : 0;
  }
  if (in) {
    int pos;
    if (!empty) {
      pos = 0;
      for (int j = 0; j < kc; ++j) pos += (j != i && head_better(s_ok[j], s_id[j], mok, mid)) ? 1 : 0;
      if (pos == kout - 1) s_kth = unordered_f32(mok);
    } else {
      pos = nvalid + empties_before;
    }
    out_key[(long long)q * kc + pos] = empty ? -CUDART_INF_F : unordered_f32(mok);
    out_idx[(long long)q * kc + pos] = empty ? -1 : mid;
    if (pos < kout) {
      // same conventions as merge_lists_warp: missing results are id -1, +inf (L2) / -inf (IP), label 0, key -inf
      const long long o = (long long)q * kout + pos;
      const float kv = unordered_f32(mok);
      if (fin_dist) fin_dist[o] = empty ? (L2 ? CUDART_INF_F : -CUDART_INF_F) : (L2 ? fmaxf(0.f, qnorm[q] - kv) : kv);
      if (fin_idx) fin_idx[o] = empty ? -1ll : mid + id_offset;
      if (fin_lbl) fin_lbl[o] = (!empty && labels) ? __ldg(labels + mid) : 0.f;
      if (fin_key) fin_key[o] = empty ? -CUDART_INF_F : kv;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float approx_worst = CUDART_INF_F;
    for (int j = 0; j < kc; ++j)
      if (cand_idx[(long long)q * kc + j] >= 0) approx_worst = fminf(approx_worst, cand_key[(long long)q * kc + j]);
    const float kth_key = s_kth;
    const bool has = nvalid >= kout;
    // One-term pass with measured residuals (qres != null): q.y - q_hi.y_hi = (q - q_hi).y + q_hi.(y - y_hi), so
    // |error| <= |q - q_hi| max|y| + |q_hi| max|y - y_hi| (Cauchy-Schwarz with the ACTUAL rounding residuals, ~2.3x below
    // the worst case 2 * 2^-9 |q||y| that `eps` encodes) + the accumulation term; 1.001 covers the fp32 rounding of the
    // norms themselves.
    const float qn = sqrtf(qnorm[q]), ymax = sqrtf(*ynorm_max);
    float bound = eps * qn * ymax;
    if (qres) {
      const float qr = sqrtf(qres[q]);
      bound = fminf(bound, 1.001f * (qr * ymax + (qn + qr) * sqrtf(*yres_max)) + accum_eps * qn * ymax);
    }
    bound *= (L2 ? 2.0f : 1.0f);
    // L2 keys taken from the accumulator (norm slice) also carry the accumulation rounding of the -|y|^2 parts
    if (L2) bound += 4.0f * 1.1920928955078125e-07f * (*ynorm_max);
    const bool ok = (nvalid >= ntotal) || (!(sat && sat[q]) && nvalid == kc && has && kth_key > approx_worst + bound);
    // uncertified queries are listed by their ORIGINAL query index when the batch is itself a compacted sub-batch
    if (!ok) uncert_list[atomicAdd(uncert_count, 1)] = qmap ? qmap[q] : q;
  }
  __syncthreads();     // the shared arrays are reused by the next query of this block
  }
}

// Exact re-rank for LARGE candidate sets (the k > 104 searches of fp32 stores: tensor-core split-precision keys, dense
// select of kc = k + slack candidates, then this kernel).  One block per query: the warps re-score the kc candidates
// exactly in fp32 against the master rows (same arithmetic as rerank_exact_kernel and the exact CUDA-core scorer), the
// block sorts the exact (key, id) pairs (bitonic, shared memory, key desc / id asc), evaluates the same certificate
//     exact_key[kout - 1] > worst approximate candidate key + B
// and writes the first kout in FINAL form (distance or merge key, global id, label).  Uncertified queries are appended
// to uncert_list for the exact CUDA-core search.   Dynamic smem: next_pow2(kc) * 8 bytes.
constexpr int RERANK_LARGE_THREADS = 256;
template <bool L2, typename IdxT>
__global__ void __launch_bounds__(RERANK_LARGE_THREADS) rerank_exact_large_kernel(
    const IdxT* __restrict__ cand_idx, const float* __restrict__ cand_key, int Q, int kc, int kout,
    const float* __restrict__ qf, const float* __restrict__ master, const float* __restrict__ ynorm, int D, float eps,
    const float* __restrict__ qnorm, const float* __restrict__ ynorm_max, long long ntotal, long long id_offset,
    const float* __restrict__ labels, float* __restrict__ out_dist, long long* __restrict__ out_idx,
    float* __restrict__ out_lbl, float* __restrict__ out_key, int* __restrict__ uncert_list,
    int* __restrict__ uncert_count) {
  extern __shared__ __align__(16) unsigned long long rl_buf[];
  __shared__ float s_worst[RERANK_LARGE_THREADS / 32];
  __shared__ int s_nvalid[RERANK_LARGE_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int q = blockIdx.x;
  if (q >= Q) return;
  const float* qr = qf + (long long)q * D;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(qr) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(master) & 15) == 0);
  float worst = CUDART_INF_F;
  int nvalid = 0;
  for (int j = w; j < kc; j += RERANK_LARGE_THREADS / 32) {
    const long long id = (long long)cand_idx[(long long)q * kc + j];
    if (id < 0) { if (lane == 0) rl_buf[j] = 0ull; continue; }
    worst = fminf(worst, cand_key[(long long)q * kc + j]);
    ++nvalid;
    const float* yr = master + id * (long long)D;
    float s = 0.f;
    if (vec4) {
      const float4* q4 = reinterpret_cast<const float4*>(qr);
      const float4* y4 = reinterpret_cast<const float4*>(yr);
      const int n4 = D >> 2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int c = lane;
      for (; c + 96 < n4; c += 128) {
        const float4 y0 = __ldg(y4 + c), y1 = __ldg(y4 + c + 32), y2 = __ldg(y4 + c + 64), y3 = __ldg(y4 + c + 96);
        const float4 x0 = q4[c], x1 = q4[c + 32], x2 = q4[c + 64], x3 = q4[c + 96];
        a0 = fmaf(x0.x, y0.x, a0); a0 = fmaf(x0.y, y0.y, a0); a0 = fmaf(x0.z, y0.z, a0); a0 = fmaf(x0.w, y0.w, a0);
        a1 = fmaf(x1.x, y1.x, a1); a1 = fmaf(x1.y, y1.y, a1); a1 = fmaf(x1.z, y1.z, a1); a1 = fmaf(x1.w, y1.w, a1);
        a2 = fmaf(x2.x, y2.x, a2); a2 = fmaf(x2.y, y2.y, a2); a2 = fmaf(x2.z, y2.z, a2); a2 = fmaf(x2.w, y2.w, a2);
        a3 = fmaf(x3.x, y3.x, a3); a3 = fmaf(x3.y, y3.y, a3); a3 = fmaf(x3.z, y3.z, a3); a3 = fmaf(x3.w, y3.w, a3);
      }
      for (; c < n4; c += 32) {
        const float4 y0 = __ldg(y4 + c);
        const float4 x0 = q4[c];
        a0 = fmaf(x0.x, y0.x, a0); a0 = fmaf(x0.y, y0.y, a0); a0 = fmaf(x0.z, y0.z, a0); a0 = fmaf(x0.w, y0.w, a0);
      }
      s = (a0 + a1) + (a2 + a3);
    } else {
      for (int c = lane; c < D; c += 32) s = fmaf(qr[c], __ldg(yr + c), s);
    }
    s = warp_sum(s);
    const float key = L2 ? fmaf(2.0f, s, -ynorm[id]) : s;
    if (lane == 0) rl_buf[j] = pack_cand(key, uint32_t(id));
  }
  if (lane == 0) { s_worst[w] = worst; s_nvalid[w] = nvalid; }
  int n2 = 1;
  while (n2 < kc) n2 <<= 1;
  for (int i = kc + threadIdx.x; i < n2; i += RERANK_LARGE_THREADS) rl_buf[i] = 0ull;      // below every real element
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n2 >> 1); i += RERANK_LARGE_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = rl_buf[lo], b = rl_buf[hi];
        if ((a < b) == desc) { rl_buf[lo] = b; rl_buf[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < kout; r += RERANK_LARGE_THREADS) {
    const unsigned long long v = rl_buf[r];
    const long long o = (long long)q * kout + r;
    if (v == 0ull) {
      if (out_dist) out_dist[o] = L2 ? CUDART_INF_F : -CUDART_INF_F;
      out_idx[o] = -1;
      if (out_lbl) out_lbl[o] = 0.f;
      if (out_key) out_key[o] = -CUDART_INF_F;
    } else {
      const float kv = unordered_f32(uint32_t(v >> 32));
      const long long id = (long long)(0xFFFFFFFFu - uint32_t(v));
      if (out_dist) out_dist[o] = L2 ? fmaxf(0.f, qnorm[q] - kv) : kv;
      out_idx[o] = id + id_offset;
      if (out_lbl) out_lbl[o] = labels ? __ldg(labels + id) : 0.f;
      if (out_key) out_key[o] = kv;
    }
  }
  if (threadIdx.x == 0) {
    float aw = CUDART_INF_F;
    int nv = 0;
    for (int i = 0; i < RERANK_LARGE_THREADS / 32; ++i) { aw = fminf(aw, s_worst[i]); nv += s_nvalid[i]; }
    const unsigned long long vk = (kout >= 1 && kout <= kc) ? rl_buf[kout - 1] : 0ull;
    const bool has = vk != 0ull;
    const float kth_key = has ? unordered_f32(uint32_t(vk >> 32)) : -CUDART_INF_F;
    float bound = eps * sqrtf(qnorm[q]) * sqrtf(*ynorm_max) * (L2 ? 2.0f : 1.0f);
    if (L2) bound += 4.0f * 1.1920928955078125e-07f * (*ynorm_max);      // norm slice: see rerank_exact_kernel
    const bool ok = (nv >= ntotal) || (nv == kc && has && kth_key > aw + bound);
    if (!ok) uncert_list[atomicAdd(uncert_count, 1)] = q;
  }
}

// ---- sampled pivot for the large-k tensor-core path (see launch_tc_pivoted in radad_flat.cu) ---------------------------
// gthr[q] <- ordered(key of rank `rank` of the sample's merged list), 0 (= no bound) when the sample holds fewer rows
static __global__ void gthr_from_pivot_kernel(const float* __restrict__ sample_key, int Q, int kc, int rank,
                                       uint32_t* __restrict__ gthr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float v = sample_key[(long long)q * kc + rank];
  gthr[q] = (v > -CUDART_INF_F) ? ordered_f32(v) : 0u;
}
// flag <- 1 if any query ended with fewer than min(k, N) neighbours (the pivot was too high for it); gthr <- 0 so the
// fallback launch starts without any bound
static __global__ void check_complete_kernel(const long long* __restrict__ out_idx, int Q, int k, int need,
                                      uint32_t* __restrict__ gthr, int* __restrict__ flag) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  gthr[q] = 0u;
  if (need > 0 && out_idx[(long long)q * k + need - 1] < 0) atomicExch(flag, 1);
}

// min |y|^2 over aligned groups of 32 rows: out[g0 + i] for i < ngroups (one warp per group).  Lets the tcgen05
// epilogue reject a whole 32-column group in the L2 metric with one compare: key_j = 2 s_j - |y_j|^2 <= 2 max(s) - min.
static __global__ void ynorm_min32_kernel(const float* __restrict__ ynorm, long long g0, long long ngroups,
                                   float* __restrict__ out) {
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= ngroups) return;
  float m = ynorm[(g0 + w) * 32 + (threadIdx.x & 31)];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) out[g0 + w] = m;
}

// Norm slice of the tcgen05 scorer (score_tc.cuh, TcParams::ext): yext[row] = {p1, p2, p3, 0, 0, 0, 0, 0} with
// p1 + p2 + p3 == -|y|^2 EXACTLY (three round-to-nearest bf16 parts hold the 24 significant bits of an fp32), and the
// constant query side {1, 1, 1, 0, ...}.
static __global__ void yext_fill_kernel(const float* __restrict__ ynorm, long long m, uint4* __restrict__ yext) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float n = -ynorm[i];
  const __nv_bfloat16 p1 = __float2bfloat16_rn(n);
  const float r1 = n - __bfloat162float(p1);
  const __nv_bfloat16 p2 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(p2);
  const __nv_bfloat16 p3 = __float2bfloat16_rn(r2);
  uint4 o;
  o.x = uint32_t(__bfloat16_as_ushort(p1)) | (uint32_t(__bfloat16_as_ushort(p2)) << 16);
  o.y = uint32_t(__bfloat16_as_ushort(p3));
  o.z = 0u; o.w = 0u;
  yext[i] = o;
}
static __global__ void qext_fill_kernel(long long n, uint4* __restrict__ qext) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  qext[i] = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);       // bf16 1.0 = 0x3F80
}

// max over rows of |y|^2 (positive floats order like their bit patterns) -- feeds the re-rank certificate
static __global__ void ynorm_max_kernel(const float* __restrict__ ynorm, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, ynorm[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

// rows list[i] of src [*, D] -> dst [m, D]   (compact the uncertified queries); m_dev (optional) = count on the device
static __global__ void gather_f32_rows_kernel(const float* __restrict__ src, const int* __restrict__ list, int m, int D,
                                       float* __restrict__ dst, const int* __restrict__ m_dev = nullptr) {
  if (m_dev) m = min(m, __ldcg(m_dev));
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < m; w += nw) {
    const float* s = src + (long long)list[w] * D;
    float* d = dst + (long long)w * D;
    for (int c = (threadIdx.x & 31); c < D; c += 32) d[c] = s[c];
  }
}

// scatter rows of the fallback results back to their query slots
static __global__ void scatter_results_kernel(const int* __restrict__ list, int m, int k, const float* __restrict__ s_a,
                                       const long long* __restrict__ s_i, const float* __restrict__ s_l,
                                       float* __restrict__ d_a, long long* __restrict__ d_i,
                                       float* __restrict__ d_l, const int* __restrict__ m_dev = nullptr) {
  if (m_dev) m = min(m, __ldcg(m_dev));
  const long long total = (long long)m * k;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = int(t / k), j = int(t % k);
    const long long o = (long long)list[r] * k + j;
    d_a[o] = s_a[t];
    d_i[o] = s_i[t];
    if (d_l && s_l) d_l[o] = s_l[t];
  }
}

// ---- planning kernels of the device-sized launches (DevPlan, common.cuh) -- one thread each ----------------------------
// same cost model as the host's choose_splits (radad_flat.cu): waves * (tiles per chunk + per-unit overhead)
// (one warp: lane l evaluates S = l + 1, l + 33, ...; 32-bit arithmetic -- a single thread looping over 128 candidates with
// 64-bit divisions took 39 us)
__device__ __forceinline__ int plan_choose_splits(int nqt, int ntiles, int slots, int max_lists, int min_tiles,
                                                  float overhead, int* tpc_out) {
  const int lane = threadIdx.x & 31;
  int maxS = min(max_lists, ntiles);
  maxS = min(maxS, max(1, ntiles / max(min_tiles, 1)));
  float best = 3.0e38f;
  int bestS = 0x7FFFFFFF;
  for (int S = lane + 1; S <= maxS; S += 32) {
    const int tpc = (ntiles + S - 1) / S;
    if ((ntiles + tpc - 1) / tpc != S) continue;
    const unsigned units = unsigned(nqt) * unsigned(S);
    const unsigned waves = (units + unsigned(slots) - 1u) / unsigned(slots);
    const float cost = float(waves) * (float(tpc) + overhead);
    if (cost < best * 0.999f) { best = cost; bestS = S; }      // ascending S per lane: the host's tie rule (smallest S)
  }
  // warp arg-min; ties (within 0.1 %) go to the smaller S like the host's ascending scan
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oS = __shfl_xor_sync(0xffffffffu, bestS, o);
    if (ob < best * 0.999f || (ob <= best * 1.001f && oS < bestS && ob < 3.0e38f)) { best = fminf(ob, best); bestS = oS; }
  }
  bestS = __shfl_sync(0xffffffffu, bestS, 0);
  if (bestS == 0x7FFFFFFF) bestS = 1;
  *tpc_out = (ntiles + bestS - 1) / bestS;
  return bestS;
}
// Tensor-core stage over `*count` queries (clamped to cap): query-tile groups of 128 * cg queries, chunking chosen to
// fill `slots` CTA groups; lists per query bounded by the candidate buffer (cand_entries slots of kc entries).
static __global__ void plan_tc_kernel(const int* __restrict__ count, int cap, int cg, long long ntiles, int slots,
                                      int max_lists, int min_tiles, float overhead, long long cand_lists_cap,
                                      DevPlan* __restrict__ plan) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;           // one warp
  const int nq = min(max(*count, 0), cap);
  if (nq == 0) {
    if (threadIdx.x == 0) { plan->nq = 0; plan->nqg = 0; plan->S = 1; plan->tpc = int(ntiles); plan->num_units = 0; plan->L = TC_PLAN_LISTS; }
    return;
  }
  const int nqg = (nq + 128 * cg - 1) / (128 * cg);
  const long long lists_fit = max(1ll, cand_lists_cap / ((long long)nqg * 128 * cg) / TC_PLAN_LISTS);
  int tpc;
  const int S = plan_choose_splits(nqg, int(ntiles), slots, int(min((long long)max_lists, lists_fit)), min_tiles, overhead, &tpc);
  if (threadIdx.x == 0) {
    plan->nq = nq; plan->nqg = nqg; plan->S = S; plan->tpc = tpc; plan->num_units = nqg * S; plan->L = S * TC_PLAN_LISTS;
  }
}
// CUDA-core stage (exact fallback) over `*count` queries: 128-query tiles x chunks of 128-row tiles, 2 lists per chunk
static __global__ void plan_simt_kernel(const int* __restrict__ count, int cap, long long ntiles, int slots, int max_lists,
                                        long long cand_lists_cap, DevPlan* __restrict__ plan) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;           // one warp
  const int nq = min(max(*count, 0), cap);
  if (nq == 0) {
    if (threadIdx.x == 0) { plan->nq = 0; plan->s_nqt = 0; plan->s_S = 1; plan->s_rows = 128; plan->s_units = 0; plan->s_L = 2; }
    return;
  }
  const int nqt = (nq + 127) / 128;
  const long long lists_fit = max(1ll, cand_lists_cap / ((long long)nqt * 128) / 2);
  int tpc;
  const int S = plan_choose_splits(nqt, int(ntiles), slots, int(min((long long)max_lists, lists_fit)), 2, 2.0f, &tpc);
  if (threadIdx.x == 0) {
    plan->nq = nq; plan->s_nqt = nqt; plan->s_S = S; plan->s_rows = tpc * 128; plan->s_units = nqt * S; plan->s_L = S * 2;
  }
}

// Fused gather + merge over NVLink peer memory (multi-GPU row shards).  Each rank left its [Q][k] candidates
// (key, global id, label) in a buffer that every peer has mapped (CUDA IPC); this kernel reads list g straight from
// GPU g's memory with ordinary loads (P2P over NVLink/NVSwitch) while merging -- no all-gather, no staging copy.
// One warp per query, lane g owns list g (G <= 32).
struct PeerLists {
  const float* key[32];
  const long long* idx[32];
  const float* lbl[32];
};
static __global__ void __launch_bounds__(128) merge_peer_lists_kernel(const PeerLists P, int G, int Q, int k, int metric_l2,
                                                               const float* __restrict__ qnorm,
                                                               float* __restrict__ out_dist,
                                                               long long* __restrict__ out_idx,
                                                               float* __restrict__ out_lbl) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const long long base = (long long)q * k;
  const float* pk = nullptr; const long long* pi = nullptr; const float* pl = nullptr;
#pragma unroll
  for (int g = 0; g < 32; ++g) if (g == lane && g < G) { pk = P.key[g]; pi = P.idx[g]; pl = P.lbl[g]; }
  int ptr = 0;
  uint32_t hok = 0; long long hid = 0x7FFFFFFFFFFFFFFFll; float hkey = 0.f;
  if (pk) {
    const long long id = pi[base];
    if (id >= 0) { hkey = pk[base]; hok = ordered_f32(hkey); hid = id; }
  }
  for (int r = 0; r < k; ++r) {
    uint32_t wok = hok; long long wid = hid; int wl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ook = __shfl_xor_sync(0xffffffffu, wok, o);
      const long long oid = __shfl_xor_sync(0xffffffffu, wid, o);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (head_better(ook, oid, wok, wid) || (ook == wok && oid == wid && ol < wl)) { wok = ook; wid = oid; wl = ol; }
    }
    const long long o = base + r;
    if (wok == 0) {
      if (lane == 0) {
        out_dist[o] = metric_l2 ? CUDART_INF_F : -CUDART_INF_F;
        out_idx[o] = -1;
        if (out_lbl) out_lbl[o] = 0.f;
      }
      continue;
    }
    if (lane == wl) {
      out_dist[o] = metric_l2 ? fmaxf(0.f, qnorm[q] - hkey) : hkey;
      out_idx[o] = hid;
      if (out_lbl) out_lbl[o] = pl ? pl[base + ptr] : 0.f;
      ++ptr;
      hok = 0; hid = 0x7FFFFFFFFFFFFFFFll;
      if (ptr < k) {
        const long long id = pi[base + ptr];
        if (id >= 0) { hkey = pk[base + ptr]; hok = ordered_f32(hkey); hid = id; }
      }
    }
    __syncwarp();
  }
}

// Rank-ordered self-exclusion + first-K-survivors compaction (pipeline.py:491-520) on the device.
//   idx/dist/lbl [B][ks] best-first search results; row_code[ntotal] = int code of each row's file basename;
//   excl [ne] sorted ascending = codes to skip.  One WARP per query: lane j tests result j (row code lookup + binary
//   search of the exclusion list run for 32 results at once -- one thread per query walked them one after the other, a
//   chain of ~10 dependent loads per result: 15 us for a batch of 256), a ballot keeps the rank order, the first K
//   survivors are written; pads with id -1 / label 0 / distance NaN.
static __global__ void __launch_bounds__(128) filter_first_k_kernel(
    const long long* __restrict__ idx, const float* __restrict__ dist, const float* __restrict__ lbl, int B, int ks,
    const long long* __restrict__ row_code, long long ntotal, const long long* __restrict__ excl, int ne, int K,
    long long* __restrict__ out_idx, float* __restrict__ out_dist, float* __restrict__ out_lbl) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int n = 0;                                              // survivors so far (warp-uniform)
  for (int j0 = 0; j0 < ks && n < K; j0 += 32) {
    const int j = j0 + lane;
    const long long id = j < ks ? idx[(long long)b * ks + j] : -1;
    bool keep = id >= 0 && id < ntotal;
    if (keep && ne > 0) {
      const long long code = row_code[id];
      int lo = 0, hi = ne - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const long long v = excl[mid];
        if (v == code) { keep = false; break; }
        if (v < code) lo = mid + 1; else hi = mid - 1;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int pos = n + __popc(mask & ((1u << lane) - 1u));
    if (keep && pos < K) {
      out_idx[(long long)b * K + pos] = id;
      out_dist[(long long)b * K + pos] = dist[(long long)b * ks + j];
      out_lbl[(long long)b * K + pos] = lbl ? lbl[(long long)b * ks + j] : 0.f;
    }
    n += __popc(mask);
  }
  for (int p = min(n, K) + lane; p < K; p += 32) {
    out_idx[(long long)b * K + p] = -1;
    out_dist[(long long)b * K + p] = __int_as_float(0x7FC00000);   // NaN (pipeline.py:515)
    out_lbl[(long long)b * K + p] = 0.f;
  }
}

// sum of neighbour labels per query over the first kvote results (the "kNN label vote" evidence)
static __global__ void label_vote_kernel(const float* __restrict__ lbl, int Q, int k, int kvote, float* __restrict__ vote) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float s = 0.f;
  for (int j = 0; j < kvote && j < k; ++j) s += lbl[(long long)q * k + j];
  vote[q] = s;
}

}  // namespace rdb
