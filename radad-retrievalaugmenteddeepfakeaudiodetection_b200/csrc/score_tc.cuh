// Kernel 2 -- score + select on tcgen05 tensor cores (the headline path).
//
// S = Q[nq, D] x Y[N, D]^T is never written to HBM.  A persistent, warp-specialised CTA owns one 128-query
// tile (the 128 TMEM lanes) and walks a contiguous chunk of database rows in 256-row tiles:
//
//   warp 0   TMA producer : K-slices (64 bf16 = one 128-byte swizzle atom) of the query tile and the DB tile
//                           into a 4-stage shared-memory ring (mbarrier full/empty).
//   warp 1   MMA issuer   : one elected thread issues tcgen05.mma.kind::f16 (M=128, N=256, K=16), fp32
//                           accumulators double-buffered in TMEM (2 x 256 of 512 columns).
//   warps 2-9 epilogue    : tcgen05.ld 32x32b -> each thread owns ONE query row x one 128-column half, applies the L2 fix-up
//                           (key = 2 q.y - |y|^2), threshold-filters against its running k-th best and inserts
//                           the rare survivors into a register-resident sorted list.  MMA of tile t+1
//                           overlaps selection of tile t.
//
// Work units are (query tile, DB chunk) pairs, ordered chunk-major so that the CTAs running concurrently
// stream the SAME database tiles (served from L2 after the first fetch) against different query tiles.
// Each unit writes its [128, kout] partial result; merge.cuh folds the S partial lists per query.
//
// Split-precision mode (nterms = 3) scores fp32 data as q_lo.y_hi + q_hi.y_lo + q_hi.y_hi with
// hi = bf16(x), lo = bf16(x - hi): relative error <= 3 * 2^-18 |q||y| before the exact fp32 re-rank.
//
// Reference being replaced: faiss GpuIndexFlat::search reached from vector_database.py:181.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace rdb {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;
// CG = CTAs per MMA (tcgen05 cta_group): 1 = one CTA owns a 128-query tile and stages whole 256-row DB tiles;
// 2 = a CTA pair (the two SMs of a TPC) runs M = 256 MMAs: each CTA stages its own 128 queries plus HALF of the DB
// tile (the tensor cores read the other half from the peer SM), i.e. 32 KB instead of 48 KB per K-slice per SM:
// a third less L2->SM operand traffic and a 6-deep instead of 4-deep ring in the same shared memory.
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
template <int CG> struct TcCfg {
  static constexpr int STAGES = (CG == 2) ? 6 : 4;      // ring slots (6 x 32 KB / 4 x 48 KB; 4..7 measured equal for the pair)
  static constexpr int B_ROWS = TC_BN / CG;
  static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr size_t SMEM = size_t(STAGES) * STAGE_BYTES + 256 + 1024;
  // Query-stationary form (D <= 256, one MMA term): the whole 128 x D query tile (<= 4 K-slices = 64 KB) is loaded
  // ONCE per work unit and stays in shared memory; the ring then holds database slices only (twice as many stages in
  // the same shared memory).  A third less L2->SM operand traffic and half the TMA / mbarrier operations per tile --
  // what matters at D = 256, where a tile is only 16 MMAs long.
  static constexpr int ASTAT_MAX_KS = 4;
  static constexpr int ASTAT_A_BYTES = ASTAT_MAX_KS * TC_A_BYTES;
  static constexpr int ASTAT_STAGES = (STAGES * STAGE_BYTES - ASTAT_A_BYTES) / B_BYTES;
  static constexpr int MAX_STAGES = ASTAT_STAGES > STAGES ? ASTAT_STAGES : STAGES;
  static_assert(MAX_STAGES * 16 + 6 * 8 + 8 <= 256, "barrier block");
};
constexpr int TC_EPI_WARPS = 8;        // two per TMEM lane quarter: each takes one half of the 256 columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_LISTS = 2;            // candidate lists per (query, chunk): one per column half
constexpr int TC_TMEM_COLS = 512;
template <int CG> constexpr size_t tc_smem_bytes() { return TcCfg<CG>::SMEM; }

struct TcParams {
  CUtensorMap tmap_q[2];  // [0] = hi, [1] = lo   bf16/f16 [nq, D], box {64, 128}, SWIZZLE_128B
  CUtensorMap tmap_y[2];  // [0] = hi, [1] = lo   bf16/f16 [N,  D], box {64, 256 / CG}, SWIZZLE_128B
  // Norm slice (L2 metric on bf16 operands, `ext` != 0): one extra K-slice per tile whose first three columns hold
  // -|y|^2 split into three bf16 parts on the database side and 1, 1, 1 on the query side, while the queries are staged
  // as 2 q.  The accumulator then IS the L2 key 2 q.y - |y|^2 (fp32 accumulation, the split is exact), and the epilogue
  // is the inner-product one: no per-column fix-up, no norm loads, no coarse filter.  Both maps are [rows, 8] arrays
  // read with the regular {64, rows} box: columns 8..63 are out of bounds and zero-filled by TMA (16 bytes of traffic
  // per row); only the first K = 16 MMA of the slice is issued.
  CUtensorMap tmap_qx, tmap_yx;
  int ext;
  const float* ynorm;     // [N] |y|^2 (L2 only)
  const float* ynmin32;   // [N / 32] min |y|^2 per aligned 32-row group (L2 only): coarse filter
  float* cand_key;        // [nq][S * TC_LISTS][kout]
  int* cand_idx;          // [nq][S * TC_LISTS][kout]  local row ids, -1 = empty
  uint32_t* gthr;         // [nq] shared lower bound of the global k-th best key (ordered uint, 0 = none), or null
  int share2;             // two-list cover (SelectSmall<16, 2>): gthr is [nq][2], see common.cuh
  int nq, N, D;
  int nqt /* query-tile groups of 128 * CG queries */, S, tiles_per_chunk, ntiles, kout, num_units, nterms;
  uint32_t idesc;
  int dbg;                  // RDB_PROFILING builds only (option "tc_debug"): 1 = skip the selection work (results invalid)
  uint64_t hint_q, hint_y;  // TMA L2 eviction-priority hints for the query / database operand
  // Lock-step window (see below): progress counters [slot][sync_span][sync_groups], or null = off
  uint32_t* sync;
  int sync_groups, sync_window, sync_span, sync_spins;   // sync_span: chunks one slot can touch;   // sync_spins: polls (~1 us each) before a producer gives lock-step up
  uint32_t* sync_broken;    // set by the first producer that gives up: nobody waits any more in this launch
  int tile_step;            // 1 = every DB tile; > 1: strided sample pass (tile index t stands for tile t * tile_step)
  int nstages;              // ring slots used (<= TcCfg::STAGES; option "tc_stages")
  int astat;                // query-stationary form (host: nterms == 1 && D <= 256)
  const int* run_if;        // fallback launch: all CTAs exit at once unless *run_if != 0 (null = always run)
  const DevPlan* plan;      // device-sized launch (see DevPlan): nq / nqt / S / tiles_per_chunk / num_units come from here
  // k > 128 path (SelectDump): the launch covers rows [row_base, N) only (row_base a multiple of 256; tile t stands
  // for rows row_base + 256 t ...) and writes key(q, row) to dump[q * dump_pitch + row - row_base]
  int row_base;
  float* dump;
  long long dump_pitch;
};

// ---- lock-step window ----------------------------------------------------------------------------------------------
// The CTAs that run the same DB chunk in the same scheduling slot (the j-th unit of every CTA) stream the SAME tiles
// against different query tiles; one HBM fetch serves all of them only while they stay within an L2-sized window of
// each other.  Left alone they drift apart (ncu, C3: 19.5 database volumes of DRAM reads per launch instead of the
// 3.5 the schedule needs).  The TMA producers therefore keep a sliding window: a producer bumps counter[g] after
// issuing group g (TC_SYNC_GS tiles) and does not start group g + W before all members of its (slot, chunk) bumped
// counter[g].  It is a performance hint only -- relaxed atomics, the counter for the next check is prefetched one
// group ahead (no stall for CTAs that are not ahead), and the wait is bounded (a member that is not resident, e.g.
// on a GPU shared with another kernel, cannot deadlock the others: the waiter gives up lock-step for that unit).
constexpr int TC_SYNC_GS = 8;
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_u32(uint32_t* p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// value of element j (dynamic) of a register array, as a 31-select tree (keeps the array in registers)
__device__ __forceinline__ float sel32(const float (&v)[32], int j) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
  float c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
  const float d0 = (j & 8) ? c[1] : c[0];
  const float d1 = (j & 8) ? c[3] : c[2];
  return (j & 16) ? d1 : d0;
}

__device__ __forceinline__ void tc_load_yn(float4 (&y)[8], const float* __restrict__ ynorm, int col0) {
  const float4* yn4 = reinterpret_cast<const float4*>(ynorm + col0);  // col0 % 32 == 0; ynorm is padded to tiles
#pragma unroll
  for (int g = 0; g < 8; ++g) y[g] = __ldg(yn4 + g);
}

// One 32-column group of one query row.  L2 metric: key_j = 2 s_j - |y_j|^2.  The exact fix-up (8 x 128-bit norm loads
// + 32 FFMA) is skipped in the common case: with ynmin = min |y|^2 over the group (precomputed at ingest),
// key_j <= 2 max_j(s_j) - ynmin, so a group whose raw maximum cannot reach the admission threshold is rejected with
// one FFMA and one compare -- the same cost as the inner-product metric.  The filter is conservative, hence exact.
template <class Sel, bool L2>
__device__ __forceinline__ void tc_process32(uint32_t (&r)[32], Sel& sel, const float* __restrict__ ynorm, float ynmin,
                                             int col0 /*global row id of column 0 of this group*/, int nvalid) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if constexpr (Sel::kDump) {
    // k > 128: no selection -- this thread's 32 keys go to its query's line of the dense key buffer (128 bytes)
    if (sel.row != nullptr && nvalid > 0) {
      if (L2) {
        float4 y[8];
        tc_load_yn(y, ynorm, col0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          v[4 * g + 0] = fmaf(2.0f, v[4 * g + 0], -y[g].x);
          v[4 * g + 1] = fmaf(2.0f, v[4 * g + 1], -y[g].y);
          v[4 * g + 2] = fmaf(2.0f, v[4 * g + 2], -y[g].z);
          v[4 * g + 3] = fmaf(2.0f, v[4 * g + 3], -y[g].w);
        }
      }
      float* dst = sel.row + col0;
      if (nvalid >= 32) {
        // four 256-bit stores (sm_100: STG.256): every store fills whole 32-byte sectors of this query's line
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * g), "f"(v[8 * g]),
                       "f"(v[8 * g + 1]), "f"(v[8 * g + 2]), "f"(v[8 * g + 3]), "f"(v[8 * g + 4]), "f"(v[8 * g + 5]),
                       "f"(v[8 * g + 6]), "f"(v[8 * g + 7])
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = v[j];
      }
    }
    return;
  }
  const float worst = sel.threshold();
  // fast path: a depth-5 max tree (31 independent FMNMX) decides whether ANY of the 32 columns can survive;
  // in steady state almost no group does, so the per-element compare/mask work below is skipped entirely.
  float m[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) m[j] = fmaxf(v[2 * j], v[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[2 * j], m[2 * j + 1]);
  float oct[4];                                   // maxima of columns [8o, 8o + 8)
#pragma unroll
  for (int o = 0; o < 4; ++o) oct[o] = fmaxf(m[2 * o], m[2 * o + 1]);
  float mx = fmaxf(fmaxf(oct[0], oct[1]), fmaxf(oct[2], oct[3]));
  if (L2) {
    if (nvalid >= 32 && !(fmaf(2.0f, mx, -ynmin) > worst)) {   // no column of this group can beat the threshold
      __syncwarp();
      sel.end_group(32);
      return;
    }
    // rare: exact keys for the whole group, then the common logic below on the fixed-up values
    float4 y[8];
    tc_load_yn(y, ynorm, col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      v[4 * g + 0] = fmaf(2.0f, v[4 * g + 0], -y[g].x);
      v[4 * g + 1] = fmaf(2.0f, v[4 * g + 1], -y[g].y);
      v[4 * g + 2] = fmaf(2.0f, v[4 * g + 2], -y[g].z);
      v[4 * g + 3] = fmaf(2.0f, v[4 * g + 3], -y[g].w);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float a = fmaxf(fmaxf(v[8 * o], v[8 * o + 1]), fmaxf(v[8 * o + 2], v[8 * o + 3]));
      float b = fmaxf(fmaxf(v[8 * o + 4], v[8 * o + 5]), fmaxf(v[8 * o + 6], v[8 * o + 7]));
      oct[o] = fmaxf(a, b);
    }
    mx = fmaxf(fmaxf(oct[0], oct[1]), fmaxf(oct[2], oct[3]));
  }
  if (nvalid < 32) {
    // ragged last tile (rare): generic 32-wide mask
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) mask |= (v[j] > worst) ? (1u << j) : 0u;
    mask &= (nvalid <= 0) ? 0u : (0xFFFFFFFFu >> (32 - nvalid));
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      sel.offer(sel32(v, j), col0 + j);
    }
  } else if (mx > worst) {
    // slow path, kept warp-convergent and narrow: lanes walk their hot octets (usually one); the octet's 8 values
    // are fetched with a 2-level select, reduced to an 8-bit mask, and survivors go through ONE offer site.
    uint32_t hot = (oct[0] > worst ? 1u : 0u) | (oct[1] > worst ? 2u : 0u) | (oct[2] > worst ? 4u : 0u) |
                   (oct[3] > worst ? 8u : 0u);
    while (hot) {
      const int o = __ffs(hot) - 1;
      hot &= hot - 1;
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = (o & 1) ? v[8 + j] : v[j];
        const float b = (o & 1) ? v[24 + j] : v[16 + j];
        w[j] = (o & 2) ? b : a;
      }
      const float th = sel.threshold();
      uint32_t m8 = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) m8 |= (w[j] > th) ? (1u << j) : 0u;
      while (m8) {
        const int j = __ffs(m8) - 1;
        m8 &= m8 - 1;
        const float x01 = (j & 1) ? w[1] : w[0], x23 = (j & 1) ? w[3] : w[2];
        const float x45 = (j & 1) ? w[5] : w[4], x67 = (j & 1) ? w[7] : w[6];
        const float xa = (j & 2) ? x23 : x01, xb = (j & 2) ? x67 : x45;
        sel.offer((j & 4) ? xb : xa, col0 + 8 * o + j);
      }
    }
  }
  __syncwarp();
  sel.end_group(32);
}

template <class Sel, bool L2, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1) score_select_tc_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<CG>;
  constexpr int STAGES = Cfg::STAGES;
  if (p.run_if && __ldcg(p.run_if) == 0) return;     // uniform across the grid (and both CTAs of a pair)
  // schedule: from the host, or (device-sized launches) from the plan a planning kernel wrote -- uniform either way
  int s_nq = p.nq, s_nqt = p.nqt, s_S = p.S, s_tpc = p.tiles_per_chunk, s_units = p.num_units;
  if (p.plan) {
    s_nq = __ldcg(&p.plan->nq); s_nqt = __ldcg(&p.plan->nqg); s_S = __ldcg(&p.plan->S); s_tpc = __ldcg(&p.plan->tpc);
    s_units = __ldcg(&p.plan->num_units);
    if (s_nq <= 0) return;
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const bool astat = p.astat != 0;
  const int nst = astat ? Cfg::ASTAT_STAGES : min(STAGES, p.nstages);  // ring depth
  const int ring_stride = astat ? Cfg::B_BYTES : Cfg::STAGE_BYTES;    // bytes per ring slot
  uint8_t* ring = smem + (astat ? Cfg::ASTAT_A_BYTES : 0);            // slot s: [A slice |] B slice
  const int b_off = astat ? 0 : TC_A_BYTES;                           // offset of the B slice inside a slot
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* afull_bar = tempty_bar + 2;                               // query tile resident (astat)
  uint64_t* aempty_bar = afull_bar + 1;                               // query tile no longer read by any MMA (astat)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CG == 2: the two CTAs of a cluster form one MMA group; `group` walks the work units, `rank` picks the query
  // tile (A half) and the database half (B half) this CTA stages.  Rank 0 is the leader (issues MMA, owns full_bar).
  const int rank = (CG == 2) ? int(cluster_ctarank()) : 0;
  const int group = int(blockIdx.x) / CG;
  const int ngroups = int(gridDim.x) / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q[0]);
    tma_prefetch_desc(&p.tmap_y[0]);
    if (p.ext) { tma_prefetch_desc(&p.tmap_qx); tma_prefetch_desc(&p.tmap_yx); }
    for (int s = 0; s < Cfg::MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], TC_EPI_WARPS * CG); }
    mbar_init(afull_bar, 1); mbar_init(aempty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair<TC_TMEM_COLS>(tmem_ptr);
    else tmem_alloc<TC_TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int nks = (p.D + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      int slot = 0;
      for (int unit = group; unit < s_units; unit += ngroups, ++slot) {
        const int qtile = (unit % s_nqt) * CG + rank, chunk = unit / s_nqt;
        const int t0 = chunk * s_tpc, t1 = min(p.ntiles, t0 + s_tpc);
        // lock-step window: members = units of this slot that run this chunk (a slot of ngroups consecutive units
        // touches at most sync_span = ceil(ngroups / nqt) + 1 chunks)
        const int u_lo = max(slot * ngroups, chunk * s_nqt);
        const int u_hi = min(min((slot + 1) * ngroups, (chunk + 1) * s_nqt), s_units);
        const uint32_t members = uint32_t(u_hi - u_lo);
        const bool counted = p.sync != nullptr && rank == 0 && members > 1;
        bool waiting = counted && ld_relaxed_u32(p.sync_broken) == 0u;
        uint32_t* ctr = p.sync + (size_t(slot) * p.sync_span + size_t(chunk - (slot * ngroups) / s_nqt)) * p.sync_groups;
        uint32_t seen = 0;
        if (astat) {
          // the unit's query tile: loaded once, after every MMA of the previous unit has retired
          mbar_wait(aempty_bar, aphase ^ 1);
          if (CG == 2) {
            if (rank == 0) mbar_expect_tx(afull_bar, 2 * (nks + p.ext) * TC_A_BYTES);
            for (int ks = 0; ks < nks; ++ks)
              tma_load_2d_pair(smem + ks * TC_A_BYTES, &p.tmap_q[0], afull_bar, ks * TC_BK, qtile * TC_BM, p.hint_q);
            if (p.ext) tma_load_2d_pair(smem + nks * TC_A_BYTES, &p.tmap_qx, afull_bar, 0, qtile * TC_BM, p.hint_q);
          } else {
            mbar_expect_tx(afull_bar, (nks + p.ext) * TC_A_BYTES);
            for (int ks = 0; ks < nks; ++ks)
              tma_load_2d(smem + ks * TC_A_BYTES, &p.tmap_q[0], afull_bar, ks * TC_BK, qtile * TC_BM, p.hint_q);
            if (p.ext) tma_load_2d(smem + nks * TC_A_BYTES, &p.tmap_qx, afull_bar, 0, qtile * TC_BM, p.hint_q);
          }
          aphase ^= 1;
        }
        for (int t = t0; t < t1; ++t) {
          if (counted) {
            const int tau = t - t0, g = tau / TC_SYNC_GS;
            if (tau % TC_SYNC_GS == 0) {
              if (waiting && g >= p.sync_window && seen < members) {
                int spins = 0;
                while ((seen = ld_relaxed_u32(ctr + g - p.sync_window)) < members) {
                  // patience is long (a producer that is ahead SHOULD wait: that is what re-aligns the wave); it only
                  // ends when a member evidently is not running (GPU shared with another kernel), and then for everybody
                  if (++spins > p.sync_spins || ld_relaxed_u32(p.sync_broken) != 0u) {
                    waiting = false;
                    red_add_u32(p.sync_broken);
                    break;
                  }
                  __nanosleep(500);
                }
              }
              if (g + 1 >= p.sync_window) seen = ld_relaxed_u32(ctr + g + 1 - p.sync_window);   // for the next check
            }
          }
          for (int term = 0; term <= p.nterms; ++term) {
            // nterms == 1: (hi, hi).  nterms == 3: (lo, hi), (hi, lo), (hi, hi) -- small terms first.  term == nterms:
            // the norm slice (ext), one slice from the [rows, 8] side arrays.
            const bool xs = term == p.nterms;
            if (xs && !p.ext) break;
            const int qsel = (p.nterms == 3 && term == 0) ? 1 : 0;
            const int ysel = (p.nterms == 3 && term == 1) ? 1 : 0;
            const CUtensorMap* mq = xs ? &p.tmap_qx : &p.tmap_q[qsel];
            const CUtensorMap* my = xs ? &p.tmap_yx : &p.tmap_y[ysel];
            const int nsl = xs ? 1 : nks;
            for (int ks = 0; ks < nsl; ++ks) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = ring + stage * ring_stride;
              const int slot_bytes = astat ? Cfg::B_BYTES : Cfg::STAGE_BYTES;
              if (CG == 2) {
                // both CTAs' bytes complete on the leader's barrier; the leader posts the expectation for both
                if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * slot_bytes);
                if (!astat) tma_load_2d_pair(sa, mq, &full_bar[stage], ks * TC_BK, qtile * TC_BM, p.hint_q);
                tma_load_2d_pair(sa + b_off, my, &full_bar[stage], ks * TC_BK,
                                 p.row_base + t * p.tile_step * TC_BN + rank * Cfg::B_ROWS, p.hint_y);
              } else {
                mbar_expect_tx(&full_bar[stage], slot_bytes);
                if (!astat) tma_load_2d(sa, mq, &full_bar[stage], ks * TC_BK, qtile * TC_BM, p.hint_q);
                tma_load_2d(sa + b_off, my, &full_bar[stage], ks * TC_BK, p.row_base + t * p.tile_step * TC_BN,
                            p.hint_y);
              }
              if (++stage == nst) { stage = 0; phase ^= 1; }
            }
          }
          if (counted && (t - t0) % TC_SYNC_GS == TC_SYNC_GS - 1) red_add_u32(ctr + (t - t0) / TC_SYNC_GS);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread; leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, aphase = 0;
      for (int unit = group; unit < s_units; unit += ngroups) {
        const int chunk = unit / s_nqt;
        const int t0 = chunk * s_tpc, t1 = min(p.ntiles, t0 + s_tpc);
        if (astat) { mbar_wait(afull_bar, aphase); tc_fence_after(); aphase ^= 1; }
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + uint32_t(acc * TC_BN);
          const int nslices = nks * p.nterms + p.ext;
          for (int ks = 0; ks < nslices; ++ks) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * ring_stride);
            const uint64_t da = make_sw128_kmajor_desc(astat ? smem_u32(smem + ks * TC_A_BYTES) : sa);
            const uint64_t db = make_sw128_kmajor_desc(sa + b_off);
            const int nkk = (p.ext && ks == nslices - 1) ? 1 : TC_BK / 16;   // norm slice: columns 0..2 only
#pragma unroll
            for (int kk = 0; kk < TC_BK / 16; ++kk) {
              if (kk >= nkk) break;
              // +32 bytes per K=16 step inside the 128-byte swizzle atom (descriptor address unit = 16 B)
              if (CG == 2) umma_f16_ss_pair(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), p.idesc, (ks | kk) != 0);
              else umma_f16_ss(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), p.idesc, (ks | kk) != 0);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            if (++stage == nst) { stage = 0; phase ^= 1; }
          }
          // accumulator complete -> epilogue (of both CTAs of a pair)
          if (CG == 2) umma_commit_pair(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        // every MMA that reads this unit's query tile has been issued: release it when they retire
        if (astat) { if (CG == 2) umma_commit_pair(aempty_bar); else umma_commit(aempty_bar); }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: fused top-k
    const int ew = warp & 3;                 // TMEM lane quarter this warp may access (hardware rule: warp_id % 4)
    const int half = (warp - 2) >> 2;        // which 128-column half of the tile this warp selects from
    const int row = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = group; unit < s_units; unit += ngroups) {
      const int qtile = (unit % s_nqt) * CG + rank, chunk = unit / s_nqt;
      const int t0 = chunk * s_tpc, t1 = min(p.ntiles, t0 + s_tpc);
      const long long q = (long long)qtile * TC_BM + row;
      Sel sel;
      sel.init(p.kout, (p.gthr && q < s_nq) ? p.gthr + q * Sel::kGthrWords : nullptr);
      if constexpr (Sel::kDump) sel.row = (q < s_nq) ? p.dump + q * p.dump_pitch - p.row_base : nullptr;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * TC_BN + half * (TC_BN / 2));
        const int n0 = p.row_base + t * p.tile_step * TC_BN + half * (TC_BN / 2);
        const int nvalid = p.N - n0;         // >= 128 for full tiles
#ifdef RDB_PROFILING
        if (p.dbg & 1) {      // bare main loop (selection skipped, results invalid): profiling builds only
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
          continue;
        }
#endif
        uint32_t ra[32], rb[32];
        // L2: minima of |y|^2 over this half-tile's four 32-row groups (one 128-bit load; n0 is a multiple of 128)
        float ymin[4] = {0.f, 0.f, 0.f, 0.f};
        if (L2) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.ynmin32 + (n0 >> 5)));
          ymin[0] = t4.x; ymin[1] = t4.y; ymin[2] = t4.z; ymin[3] = t4.w;
        }
        tmem_ld_32x32(taddr, ra);
#pragma unroll
        for (int c = 0; c < TC_BN / 64; c += 2) {
          tmem_ld_wait_regs(ra);
          tmem_ld_32x32(taddr + (c + 1) * 32, rb);
          tc_process32<Sel, L2>(ra, sel, p.ynorm, ymin[c], n0 + c * 32, nvalid - c * 32);
          tmem_ld_wait_regs(rb);
          if (c + 2 < TC_BN / 64) tmem_ld_32x32(taddr + (c + 2) * 32, ra);
          else {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
          }
          tc_process32<Sel, L2>(rb, sel, p.ynorm, ymin[c + 1], n0 + (c + 1) * 32, nvalid - (c + 1) * 32);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (q < s_nq) {
        const long long base = ((q * s_S + chunk) * TC_LISTS + half) * (long long)p.kout;
        sel.finalize(p.kout, p.cand_key + base, p.cand_idx + base);
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer's smem/TMEM/barriers stay alive until both are done
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair<TC_TMEM_COLS>(tmem_base);
    else tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace rdb
