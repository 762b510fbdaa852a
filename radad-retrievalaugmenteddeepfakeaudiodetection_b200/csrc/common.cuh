// Common device helpers for the RADAD flat-search kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math_constants.h>

namespace rdb {

// ----------------------------------------------------------------------------------------------
// Candidate ordering.  Everywhere inside the kernels "key" is a float where LARGER IS BETTER:
//   IP metric : key = <q, y>
//   L2 metric : key = 2<q, y> - |y|^2          (distance = |q|^2 - key, clamped at 0)
// Ties on key are broken by the LOWEST row id (the same rule the test oracle uses).
// A (key, id) pair is packed into one uint64 whose unsigned order is exactly that rule.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_f32(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t pack_cand(float key, uint32_t id_lo_is_better) {
  // id_lo_is_better: smaller wins -> store complement
  return (uint64_t(ordered_f32(key)) << 32) | uint64_t(0xFFFFFFFFu - id_lo_is_better);
}

// ----------------------------------------------------------------------------------------------
// Per-thread running top-KT list kept in registers, sorted best-first (descending key).
// Elements arrive in ascending id order, and only when strictly greater than the current
// worst, so equal keys keep the lower id and stay in ascending-id order.
// ----------------------------------------------------------------------------------------------
template <int KT>
struct TopK {
  float key[KT];
  int   idx[KT];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KT; ++j) { key[j] = -CUDART_INF_F; idx[j] = -1; }
  }
  __device__ __forceinline__ float worst() const { return key[KT - 1]; }
  __device__ __forceinline__ void insert(float v, int id) {
    key[KT - 1] = v; idx[KT - 1] = id;
#pragma unroll
    for (int j = KT - 1; j > 0; --j) {
      const bool sw = key[j] > key[j - 1];
      const float ka = key[j - 1], kb = key[j];
      const int   ia = idx[j - 1], ib = idx[j];
      key[j - 1] = sw ? kb : ka;  key[j] = sw ? ka : kb;
      idx[j - 1] = sw ? ib : ia;  idx[j] = sw ? ia : ib;
    }
  }
};

__device__ __forceinline__ float unordered_f32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// Selector policy used by the tensor-core epilogue for k <= KT (register-resident list).
// Cross-unit threshold sharing.  gthr[q] holds (as an ordered uint, 0 = none) a lower bound of the GLOBAL k-th best
// key of query q over everything any work unit has seen so far: a unit's local k-th best can only be <= the global
// one, so publishing it with atomicMax is always safe.  Units run chunk-major, so a unit usually starts with the
// bound left by the previous chunks and admits ~k/j new elements instead of k*ln(n/k).  Elements EQUAL to the bound
// must still be admitted (the bound may come from rows with higher ids), hence floor = the next float below it.
__device__ __forceinline__ float floor_from_gthr(uint32_t g) {
  return (g >= 2u) ? unordered_f32(g - 1u) : -CUDART_INF_F;
}

// SHARE = 2 ("two-list cover", tier 1 of the certified fp32 search): the merged lists must contain the global best 2 KT
// keys while every unit keeps only KT.  gthr[q] is then a PAIR (hi word = largest, lo word = second largest bound any
// unit has published; one 64-bit CAS): the two largest published local KT-th keys come from two different units with
// disjoint rows, each holding KT rows at or above its own bound, so 2 KT rows reach the SECOND largest bound -- a lower
// bound of the global (2 KT)-th best key, and the admission floor of this mode.  A row of the global best 2 KT then
// passes every floor and can only be lost by a unit that holds KT better rows of its own: the merge flags queries where
// a full list was consumed entirely (merge_lists_*: `sat`), and those queries are not certified.
template <int KT, int SHARE = 1>
struct SelectSmall {
  static constexpr bool kDump = false;
  static constexpr int kGthrWords = SHARE;
  TopK<KT> top;
  float floor_;
  uint32_t* gq;
  __device__ __forceinline__ void init(int, uint32_t* gthr_q) {
    top.init();
    gq = gthr_q;
    floor_ = gq ? floor_from_gthr(__ldcg(gq)) : -CUDART_INF_F;      // SHARE == 2: word 0 = the second largest bound
  }
  __device__ __forceinline__ float threshold() const { return fmaxf(top.worst(), floor_); }
  __device__ __forceinline__ void offer(float v, int id) { if (v > threshold()) top.insert(v, id); }
  __device__ __forceinline__ void end_group(int) {}
  __device__ __forceinline__ void finalize(int kout, float* __restrict__ ck, int* __restrict__ ci) {
    // The k-th best key is taken as a running minimum over the written prefix (the list is sorted best-first), NOT as
    // top.key[kout - 1]: an equality-selected element is turned into a dynamically indexed load by the compiler, which
    // gives the whole list a local-memory home that every insertion then has to keep up to date (round 1: 146 STL per
    // instantiation, 266 M local stores and 21 GB of L2 writes per C3 launch).
    float kth = CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      if (j < kout) { ck[j] = top.key[j]; ci[j] = top.idx[j]; kth = fminf(kth, top.key[j]); }
    }
    if (gq && kth > -CUDART_INF_F && kth < CUDART_INF_F) {
      if (SHARE == 1) {
        atomicMax(gq, ordered_f32(kth));
      } else {
        unsigned long long* g = reinterpret_cast<unsigned long long*>(gq);
        const unsigned long long v = ordered_f32(kth);
        unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(g);
        while (true) {
          const unsigned long long m1 = old >> 32, m2 = old & 0xFFFFFFFFull;
          unsigned long long nw;
          if (v > m1) nw = (v << 32) | m1;
          else if (v > m2) nw = (m1 << 32) | v;
          else break;
          const unsigned long long prev = atomicCAS(g, old, nw);
          if (prev == old) break;
          old = prev;
        }
      }
    }
  }
};

// Selector policy for large k (k <= 128): an append-only per-thread reservoir in LOCAL memory (L1/L2-backed)
// with a stale admission threshold, pruned warp-synchronously to the exact best k whenever any lane runs out
// of room.  The prune finds the k-th largest key exactly by 4-way bisection over the ordered-uint key space
// (<= 16 passes) and compacts in arrival order, so among equal keys the earliest arrivals (= lowest ids) stay.
// Appends are ~k*ln(n/k) per thread per unit and cost one scattered local store each.
template <int CAP>
struct SelectReservoir {
  static constexpr bool kDump = false;
  static constexpr int kGthrWords = 1;
  static constexpr int B = 16;   // entries loaded per batch: independent local-memory loads in flight per thread
  uint32_t okey[CAP];   // ordered_f32(key); 0 = never a valid key of a finite score
  int idx[CAP];
  int cnt, k;
  float thr;
  uint32_t* gq;
  __device__ __forceinline__ void init(int k_, uint32_t* gthr_q) {
    cnt = 0; k = k_; gq = gthr_q;
    thr = gq ? floor_from_gthr(__ldcg(gq)) : -CUDART_INF_F;
  }
  // tighten the admission threshold with what other units (and the other column half) have published
  __device__ __forceinline__ void refresh() { if (gq) thr = fmaxf(thr, floor_from_gthr(__ldcg(gq))); }
  __device__ __forceinline__ float threshold() const { return thr; }
  __device__ __forceinline__ void offer(float v, int id) {
    if (v > thr) { okey[cnt] = ordered_f32(v); idx[cnt] = id; ++cnt; }
  }
  __device__ __forceinline__ void load_keys(int i, uint32_t (&o)[B]) const {
#pragma unroll
    for (int j = 0; j < B; ++j) o[j] = (i + j < cnt) ? okey[i + j] : 0u;   // 0 is below every real key
  }
  // exact prune to the best k entries (no-op for lanes holding <= k); all loops are warp-convergent
  __device__ __noinline__ void prune() {
    if (cnt <= k) { refresh(); return; }
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = 0; i < cnt; i += B) {
      uint32_t o[B];
      load_keys(i, o);
#pragma unroll
      for (int j = 0; j < B; ++j) { hi = max(hi, o[j]); lo = min(lo, (i + j < cnt) ? o[j] : 0xFFFFFFFFu); }
    }
    // smallest t in [lo, hi] with count(okey > t) < k  ==  the k-th largest key.  8-way bisection: 3 bits per pass.
    const uint32_t kk = uint32_t(k);
    while (lo < hi) {
      const uint32_t span = hi - lo;
      uint32_t qv[7], c[7];
#pragma unroll
      for (int m = 0; m < 7; ++m) { qv[m] = lo + uint32_t((uint64_t(span) * uint32_t(m + 1)) >> 3); c[m] = 0; }
      for (int i = 0; i < cnt; i += B) {
        uint32_t o[B];
        load_keys(i, o);
#pragma unroll
        for (int j = 0; j < B; ++j)
#pragma unroll
          for (int m = 0; m < 7; ++m) c[m] += (o[j] > qv[m]) ? 1u : 0u;
      }
      // c[] is non-increasing in m; pick the first sub-interval whose upper end already has < k above it
      uint32_t nlo = qv[6] + 1, nhi = hi;
#pragma unroll
      for (int m = 6; m >= 0; --m)
        if (c[m] < kk) { nhi = qv[m]; nlo = (m == 0) ? lo : qv[m - 1] + 1; }
      lo = nlo; hi = nhi;
    }
    const uint32_t t = lo;
    uint32_t gt = 0;
    for (int i = 0; i < cnt; i += B) {
      uint32_t o[B];
      load_keys(i, o);
#pragma unroll
      for (int j = 0; j < B; ++j) gt += (o[j] > t) ? 1u : 0u;
    }
    int need = k - int(gt);                // entries equal to t to keep, in arrival (= ascending id) order
    int w = 0;
    for (int i = 0; i < cnt; i += B) {
      uint32_t o[B]; int id[B];
      load_keys(i, o);
#pragma unroll
      for (int j = 0; j < B; ++j) id[j] = (i + j < cnt) ? idx[i + j] : 0;
#pragma unroll
      for (int j = 0; j < B; ++j) {
        bool keep = o[j] > t;
        if (!keep && o[j] == t && i + j < cnt && need > 0) { keep = true; --need; }
        if (keep) { okey[w] = o[j]; idx[w] = id[j]; ++w; }     // w <= i + j: never overtakes the reads
      }
    }
    cnt = w;
    thr = fmaxf(thr, unordered_f32(t));
    if (gq) atomicMax(gq, t);            // t = exact k-th best of everything this thread has seen: a global lower bound
    refresh();
  }
  __device__ __forceinline__ void end_group(int room) {
    if (__any_sync(0xffffffffu, cnt > CAP - room)) prune();
  }
  // sorted best-first output (key desc, id asc): prune to <= k entries, then rank each entry by counting
  __device__ __noinline__ void finalize(int kout, float* __restrict__ ck, int* __restrict__ ci) {
    prune();
    for (int r = cnt; r < kout; ++r) { ck[r] = -CUDART_INF_F; ci[r] = -1; }
    for (int a = 0; a < cnt; ++a) {
      const uint32_t oa = okey[a];
      const int ia = idx[a];
      int rank = 0;
      for (int i = 0; i < cnt; i += B) {
        uint32_t o[B]; int id[B];
        load_keys(i, o);
#pragma unroll
        for (int j = 0; j < B; ++j) id[j] = (i + j < cnt) ? idx[i + j] : 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < B; ++j) rank += (o[j] > oa || (o[j] == oa && id[j] < ia)) ? 1 : 0;
      }
      if (rank < kout) { ck[rank] = unordered_f32(oa); ci[rank] = ia; }
    }
  }
};

// "Selector" of the k > 128 path on 16-bit stores: no selection at all -- the tensor-core epilogue writes every key of
// its query row to a dense [queries][rows] buffer in HBM (row = this query's line of it, or null for padding rows of
// the query tile) and select_dense_kernel (select_large.cuh) picks the best k afterwards.
struct SelectDump {
  static constexpr bool kDump = true;
  static constexpr int kGthrWords = 1;
  float* row;          // &dump[q][0] - row_base, so row[global row id] is this query's slot for that row
  __device__ __forceinline__ void init(int, uint32_t*) { row = nullptr; }
  __device__ __forceinline__ float threshold() const { return -CUDART_INF_F; }
  __device__ __forceinline__ void offer(float, int) {}
  __device__ __forceinline__ void end_group(int) {}
  __device__ __forceinline__ void finalize(int, float*, int*) {}
};

// ----------------------------------------------------------------------------------------------
// Device-side work plan.  The certified search of fp32 stores re-searches only the queries a cheaper pass could not
// certify; how many there are is known on the DEVICE only.  Instead of reading the count back (a host round trip per
// batch), the follow-up launches are always enqueued with grids sized for the worst case and read their real extent --
// number of queries, the chunking of the database that fills the machine for that many queries -- from a DevPlan a
// one-thread planning kernel derived from the count.  nq == 0 makes every consumer return at once.
// ----------------------------------------------------------------------------------------------
struct DevPlan {
  int nq;                         // queries of this stage
  int nqg, S, tpc, num_units, L;  // tensor-core scorer: query-tile groups, chunks, tiles per chunk, units, lists per query
  int s_nqt, s_S, s_rows, s_units, s_L;   // CUDA-core scorer: query tiles, chunks, rows per chunk, units, lists per query
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (Blackwell 5th-gen tensor cores, TMEM).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) { }
}

// 2-D tiled TMA load, completion signalled on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(cache_hint)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// L2 eviction-priority policies (same encodings CUTLASS uses for TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst  = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast   = 0x14F0000000000000ull;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; bf16/f16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of one cluster (the two SMs of a TPC) run ONE MMA of M = 256.
// Each CTA stages its own 128 query rows (A) and HALF of the database tile (B) in its own shared memory at the
// same offsets; the leader (cluster rank 0) issues the MMA, which reads B from both SMs.  All TMA completions
// land on the LEADER's mbarrier (peer bit 24 of the shared::cluster address cleared); tcgen05.commit multicasts
// its arrival to the same barrier offset in both CTAs.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                                 uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(cache_hint)
      : "memory");
}
// arrive on the LEADER CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
               : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(uint16_t(3))
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but with the destination registers as in/out operands: a true data dependency, so the compiler
// cannot schedule any use of r[] above the wait.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as [rows][64] with the
// 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups of 1024 B.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for SW128 K-major: 1)
//   bits [32,46) stride byte offset >> 4 (1024 B between 8-row groups)   bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// Instruction descriptor, kind::f16: fp32 accumulate, A/B both K-major.
//   [4,6) c_format=1 (F32)  [7,10) a_format  [10,13) b_format (0=F16, 1=BF16)
//   [15] a_major=0 (K)  [16] b_major=0 (K)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int ab_format) {
  return (1u << 4) | (uint32_t(ab_format) << 7) | (uint32_t(ab_format) << 10) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rdb
