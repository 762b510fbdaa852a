// Kernel 3 -- small-batch score + select (nq <= 4; the batch-1 streaming latency path, `predict()`:
// pipeline.py:1046-1054, app.py:278).  With one or two queries the search is a pure stream of the database
// through HBM (2*D flops per 2*D or 4*D bytes), so this kernel is built around the memory system instead of the
// tensor cores: every warp reads whole rows with 128-bit coalesced loads (rows are contiguous -> full DRAM
// pages, unlike 128-byte TMA box slices), 4 rows in flight per warp, fp32 FMA against the queries held in
// shared memory, butterfly reduction.  Arithmetic is exact fp32 for fp32 stores (no split / re-rank needed) and
// fp32-accumulated products of the stored 16-bit values otherwise.
//
// ONE launch does the whole search (the latency path is launch- and round-trip-bound once the stream itself runs at
// HBM speed): every block prepares the queries itself (normalise, round to the store dtype, |q|^2 -- the same
// arithmetic as ingest.cuh), streams its slice of the rows, and the LAST block to finish (atomic ticket) folds the
// per-block results into the final best-first lists -- distances, global ids, labels -- in the caller's buffers.
//
// Selection is a policy:
//   LIST   (k <= 32: 1 entry per lane, k <= 128: 4 per lane)  warp-distributed sorted top-k in registers (lane j holds
//          ranks [KL*j, KL*j+KL); insertion is a ballot + shuffle-shift), 16 warp lists merged per block, the
//          gridDim.x block lists merged by the last block.
//   FILTER (large k on large databases)  maintaining k = 100 sorted entries per warp costs more issue slots than the
//          stream itself (each warp sees only N / 2368 rows, so ~10 % of its rows still insert).  Instead a first
//          LIST launch over a strided SAMPLE of the rows (every m-th warp step, m = min(64, steps per warp) >= 4)
//          yields a per-query pivot (the sample's rank-r key, r = max(16, 4k / m) <= 128: ~r m >= 4k rows of the whole
//          database beat it), and the full pass merely appends the rows with key >= pivot to a per-query buffer; the
//          last block sorts those ~400-1000 candidates (bitonic, shared memory) and emits the first k.  Exactness does not depend
//          on the sample: the result is the exact top-k whenever k <= count <= capacity, and otherwise a flag makes
//          the (always enqueued, normally empty) LIST launch redo the search.
#pragma once
#include "common.cuh"
#include "ingest.cuh"
#include "merge.cuh"

namespace rdb {

constexpr int STREAM_THREADS = 512;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
// rows in flight per warp: 4 fp32 rows or 8 16-bit rows (~12 KB of loads outstanding per warp either way at D = 768)

template <typename T> struct StreamVec;
template <> struct StreamVec<float> {
  static constexpr int EPV = 4;   // elements per 128-bit load
  static constexpr int R = 4;     // rows in flight per lane group
  using Raw = float4;
  __device__ static __forceinline__ Raw load_raw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& x, float (&v)[4]) { v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
};
template <> struct StreamVec<__nv_bfloat16> {
  static constexpr int EPV = 8;
  static constexpr int R = 8;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& x, float (&v)[8]) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};
template <> struct StreamVec<__half> {
  static constexpr int EPV = 8;
  static constexpr int R = 8;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load_raw(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& x, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

// Warp-distributed sorted list, KL entries per lane: lane j holds ranks [KL*j, KL*j + KL) (key desc; equal keys in
// arrival = ascending-id order).  KL = 1 serves k <= 32, KL = 4 serves k <= 128 -- registers only, no shared memory.
template <int KL>
struct WarpList {
  float key[KL]; int idx[KL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KL; ++s) { key[s] = -CUDART_INF_F; idx[s] = -1; }
  }
  __device__ __forceinline__ float threshold(int k) const {
    const int r = k - 1;
    float v = key[0];
#pragma unroll
    for (int s = 1; s < KL; ++s) v = ((r % KL) == s) ? key[s] : v;
    return __shfl_sync(0xffffffffu, v, r / KL);
  }
  // all lanes call with the same (v, id); v > threshold(k)
  __device__ __forceinline__ void insert(float v, int id, int lane) {
    int pos = 0;                                          // entries >= v stay in front (stable)
#pragma unroll
    for (int s = 0; s < KL; ++s) pos += __popc(__ballot_sync(0xffffffffu, key[s] >= v));
    const int L = pos / KL, sl = pos % KL;
    const float ck = __shfl_up_sync(0xffffffffu, key[KL - 1], 1);   // previous lane's last entry
    const int ci = __shfl_up_sync(0xffffffffu, idx[KL - 1], 1);
    if (lane > L) {
#pragma unroll
      for (int s = KL - 1; s > 0; --s) { key[s] = key[s - 1]; idx[s] = idx[s - 1]; }
      key[0] = ck; idx[0] = ci;
    } else if (lane == L) {
#pragma unroll
      for (int s = KL - 1; s > 0; --s)
        if (s > sl) { key[s] = key[s - 1]; idx[s] = idx[s - 1]; }
#pragma unroll
      for (int s = 0; s < KL; ++s)
        if (s == sl) { key[s] = v; idx[s] = id; }
    }
  }
};



enum { STREAM_LIST1 = 0, STREAM_LIST4 = 1, STREAM_FILTER = 2 };
constexpr int STREAM_FCAP = 4096;       // FILTER: candidate slots per query
constexpr int STREAM_SAMPLE = 64;       // FILTER: the pivot pass reads every 64th warp step of the rows
constexpr int STREAM_QINLINE = 2048;    // floats of query data a launch can carry in its parameters (8 KB: 2 x 768 ... 1 x 2048)
constexpr int STREAM_PIVOT_RANK = 16;   // FILTER: minimum pivot rank in the sample (host raises it for thin samples)

struct StreamCtl {                      // device control block owned by the handle (zero-initialised once)
  unsigned int ticket;                  // blocks finished (self-resetting)
  int fallback;                         // FILTER pass could not produce the result -> the LIST launch must run
  int fcount[4];                        // FILTER: appended candidates per query (self-resetting)
  float pivot[4];                       // FILTER: per-query pivot key from the sample pass
  unsigned int next;                    // row chunks handed out beyond every warp's first one (self-resetting)
};

struct StreamParams {
  const void* Y; int ld; const float* ynorm; int N;       // stored rows [N, ld] (T), |y|^2
  const float* q_raw; int nq, D, normalize;               // raw fp32 queries [nq, D] (device)
  NpPlan np;                                              // numpy summation order of the row norm (ingest.cuh)
  int rows_per_block, kout, lpr_log2, step_mul;           // step_mul > 1: strided sample pass (every step_mul-th warp step)
  int chunk_steps;                                        // warp steps per claimed row chunk (see the main loop)
  float* cand_key; int* cand_idx;                         // LIST: block lists [nq][gridDim.x][kout]
  float* fkey; int* fidx;                                 // FILTER: candidates [nq][STREAM_FCAP]
  StreamCtl* ctl;
  int use_pivot_out;                                      // sample pass: only write ctl->pivot (rank STREAM_PIVOT_RANK)
  int run_if_fallback;                                    // fallback launch: exit at once unless ctl->fallback != 0
  int metric_l2; long long id_offset; const float* labels;
  float* out_dist; long long* out_idx; float* out_lbl; float* out_key; float* out_qnorm;
  // Latency path with HOST buffers: the queries travel inside the kernel parameters (q_raw == null: no host-to-device
  // copy), the outputs point into mapped pinned host memory, and the last thing the search does is publish `flag_seq`
  // in *host_flag (system-scope fence first) -- the host spins on it instead of a device-to-host copy + stream sync.
  unsigned int* host_flag; unsigned int flag_seq;
  unsigned long long* prof;   // RDB_PROFILING builds: [gridDim.x][8] globaltimer stamps per block (option "stream_prof"), else null
  float qin[STREAM_QINLINE];
};

#ifdef RDB_PROFILING
#define STREAM_STAMP(i)                                                                                   \
  do {                                                                                                    \
    if (p.prof && threadIdx.x == 0) {                                                                     \
      unsigned long long t_;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                              \
      p.prof[blockIdx.x * 8 + (i)] = t_;                                                                  \
    }                                                                                                     \
  } while (0)
#else
#define STREAM_STAMP(i) do {} while (0)
#endif

template <> __device__ __forceinline__ float to16<float>(float v) { return v; }
template <> __device__ __forceinline__ float from16<float>(float v) { return v; }

// Query prep of ONE query by one warp, arithmetic identical to ingest_rows_kernel (same loop order and fmaf chain):
// v = x / (|x| + 1e-12) when `normalize`; stored value = round_T(v) (16-bit stores) or v (fp32); returns the sum of
// squares of the STORED values.  dst is fp32 [ld], pad columns [D, ld) zero.
template <typename T>
__device__ __forceinline__ float stream_prep_query(const float* __restrict__ xr, int D, int ld, int normalize,
                                                   float* __restrict__ dst, int lane, const NpPlan& np, float* lv) {
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0);
  float denom = 1.0f;
  if (normalize) denom = __fsqrt_rn(np_sumsq_global<false>(xr, np, lv, lane)) + 1e-12f;   // numpy order: see NpPlan
  float acc = 0.f;
  if (vec4) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    for (int c = lane; c < (ld >> 2); c += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < (D >> 2)) {
        v = x4[c];
        if (normalize) { v.x = v.x / denom; v.y = v.y / denom; v.z = v.z / denom; v.w = v.w / denom; }
      }
      if (sizeof(T) == 2) {
        v.x = from16<T>(to16<T>(v.x)); v.y = from16<T>(to16<T>(v.y));
        v.z = from16<T>(to16<T>(v.z)); v.w = from16<T>(to16<T>(v.w));
      }
      reinterpret_cast<float4*>(dst)[c] = v;
      if (c < (D >> 2)) { acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc); }
    }
  } else {
    for (int c = lane; c < ld; c += 32) {
      float v = 0.f;
      if (c < D) { v = xr[c]; if (normalize) v = v / denom; }
      if (sizeof(T) == 2) v = from16<T>(to16<T>(v));
      dst[c] = v;
      if (c < D) acc = fmaf(v, v, acc);
    }
  }
  return warp_sum(acc);
}

// Y [N, ld] stored rows (T = fp32 master or the 16-bit store), ld multiple of EPV, columns [D, ld) zero (16-bit) --
// D itself must be a multiple of EPV for fp32 (checked by the host; otherwise another scorer is used).
template <typename T, int NQ, bool L2, int MODE>
__global__ void __launch_bounds__(STREAM_THREADS, 1) score_select_stream_kernel(const __grid_constant__ StreamParams p) {
  constexpr int KL = (MODE == STREAM_LIST4) ? 4 : 1;
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_qnorm[NQ];
  __shared__ float s_pivot[NQ];
  __shared__ unsigned int s_ticket;
  float* qs = sm;                                        // [NQ][ld]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = p.ld, nq = p.nq, kout = p.kout, N = p.N;
  const T* __restrict__ Y = reinterpret_cast<const T*>(p.Y);
  if (p.run_if_fallback && __ldcg(&p.ctl->fallback) == 0) {
    // nothing to redo: the FILTER pass (previous launch, complete) already wrote the results
    if (p.host_flag && blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *p.host_flag = p.flag_seq; }
    return;
  }

  STREAM_STAMP(0);
  // Lane groups of `lpr` lanes own rows (short rows: 8 or 16 lanes per row, so the shuffle reduction is amortised over
  // several rows per warp step).  Per step a warp takes RW = R * G consecutive rows: group g rows r0 + g*R .. + R.
  constexpr int R = StreamVec<T>::R;
  const int lpr_log2 = p.lpr_log2;
  const int lpr = 1 << lpr_log2, G = 32 >> lpr_log2;
  const int grp = lane >> lpr_log2, l = lane & (lpr - 1);
  const int RW = R * G;
  // Rows are handed out in chunks of CH warp steps (CH * RW rows, ~48 KB at D = 768): every warp owns chunk
  // blockIdx.x * 16 + warp to begin with and claims further ones from a global counter, so no SM sits idle while
  // another still has a long static slice left (ncu on the static partition: SMs active 88 % of the kernel's duration
  // while the active part already ran at 97 % of the HBM copy peak).  The next claim is issued before the current chunk
  // is processed, so its latency hides behind ~48 KB of loads.  The sample pass (step_mul > 1) visits every
  // step_mul-th step: chunk v, step j -> rows (v * CH + j) * step_mul * RW ...
  const int CH = p.chunk_steps;
  const long long rows_per_step = (long long)RW * p.step_mul;
  const int total_steps = int((N + rows_per_step - 1) / rows_per_step);
  // the first 7/8 of the steps go out in chunks of CH, the rest one step at a time: a warp streams ~2.8 GB/s, so a
  // 48 KB chunk is 17 us of work and whole chunks to the end would leave that much imbalance between the SMs
  const int big_chunks = (CH > 1) ? int(((long long)total_steps * 7 / 8) / CH) : 0;
  const int steps_big = big_chunks * CH;
  const int total_chunks = big_chunks + (total_steps - steps_big);
  const int first_chunk = blockIdx.x * STREAM_WARPS + warp;
  // ---- the first rows of this warp's first chunk start travelling HBM -> L2 while the queries are prepared
  if (first_chunk < total_chunks) {
    const long long row0 = (first_chunk < big_chunks ? (long long)first_chunk * CH : (long long)steps_big + (first_chunk - big_chunks)) * rows_per_step;
    const long long bytes = min((long long)(N - row0), (long long)2 * RW) * ld * (long long)sizeof(T);
    const char* base = reinterpret_cast<const char*>(Y + row0 * ld);
    for (long long off = (long long)lane * 128; off < bytes; off += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }
  // ---- query prep (every block, redundantly: nq * D elements)
  const float* qsrc = p.q_raw;
  if (qsrc == nullptr) {
    // queries inside the kernel parameters -> staged behind the prepared queries in shared memory
    float* stage = sm + ((NQ * ld + NQ * p.np.nleaves + 3) & ~3);     // 16-byte aligned
    for (int i = threadIdx.x; i < nq * p.D; i += STREAM_THREADS) stage[i] = p.qin[i];
    __syncthreads();
    qsrc = stage;
  }
  if (warp < NQ) {
    if (warp < nq) {
      const float n2 = stream_prep_query<T>(qsrc + (long long)warp * p.D, p.D, ld, p.normalize, qs + warp * ld, lane, p.np,
                                               sm + NQ * ld + warp * p.np.nleaves);
      if (lane == 0) {
        s_qnorm[warp] = n2;
        if (MODE == STREAM_FILTER) s_pivot[warp] = __ldcg(&p.ctl->pivot[warp]);
        if (blockIdx.x == 0 && p.out_qnorm) p.out_qnorm[warp] = n2;
      }
    } else {
      for (int c = lane; c < ld; c += 32) qs[warp * ld + c] = 0.f;
    }
  }
  __syncthreads();
  STREAM_STAMP(1);

  constexpr int EPV = StreamVec<T>::EPV;
  const int nvec = ld / EPV;                             // 128-bit vectors per row
  const int row_end = N;

  WarpList<KL> top[NQ];
  float thr[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    top[q].init();
    thr[q] = -CUDART_INF_F;
    // FILTER admits key >= pivot: compare against the next float below it
    if (MODE == STREAM_FILTER) thr[q] = (q < nq) ? floor_from_gthr(ordered_f32(s_pivot[q])) : CUDART_INF_F;
  }

  // (A manually double-buffered form of the step loop -- next slice's loads issued before the current one is consumed --
  // was measured SLOWER: fp32 0.479 -> 0.596 ms, bf16 0.272 -> 0.281 ms at C4; the compiler's own schedule of the plain
  // loop keeps more independent loads in flight than two conditional sets of R.)
  const unsigned int claim_base = gridDim.x * STREAM_WARPS;
  for (int chunk = first_chunk; chunk < total_chunks;) {
    unsigned int nxt = 0;
    if (lane == 0) nxt = atomicAdd(&p.ctl->next, 1u);    // claim for the NEXT round; consumed after this chunk
    const int s_begin = chunk < big_chunks ? chunk * CH : steps_big + (chunk - big_chunks);
    const int s_end = chunk < big_chunks ? s_begin + CH : s_begin + 1;
  for (int st = s_begin; st < s_end; ++st) {
    const int r0 = int((long long)st * rows_per_step);
    float acc[R][NQ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[r][q] = 0.f;
    // U slices of R rows are loaded before any of them is consumed: R * U independent 128-bit loads per lane (8 KB per
    // warp with one or two queries).  Left to the compiler the slice loop was unrolled or not depending on the code
    // around it (fp32 C4: 12 loads in flight per lane in one build, 4 in the next: 0.48 vs 0.56 ms).
    constexpr int U = ((NQ <= 2) ? 16 : 8) / R;
    for (int c = l; c < nvec; c += lpr * U) {
      typename StreamVec<T>::Raw raw[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cc = c + u * lpr;
        if (cc < nvec) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int row = min(r0 + grp * R + r, row_end - 1);   // clamp: duplicates are discarded below
            raw[u][r] = StreamVec<T>::load_raw(Y + (long long)row * ld + cc * EPV);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cc = c + u * lpr;
        if (cc < nvec) {
          float qv[NQ][EPV];
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int e = 0; e < EPV; e += 4) {
              const float4 t4 = *reinterpret_cast<const float4*>(qs + q * ld + cc * EPV + e);
              qv[q][e] = t4.x; qv[q][e + 1] = t4.y; qv[q][e + 2] = t4.z; qv[q][e + 3] = t4.w;
            }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float y[EPV];
            StreamVec<T>::unpack(raw[u][r], y);
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
              for (int e = 0; e < EPV; ++e) acc[r][q] = fmaf(qv[q][e], y[e], acc[r][q]);
          }
        }
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[r][q] += __shfl_xor_sync(0xffffffffu, acc[r][q], o);
    // keys of this lane group's R rows (lane l == 0 of each group holds the full sums after the butterfly: all do)
    bool any = false;
    float keyv[R][NQ];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r0 + grp * R + r;
      float yn = 0.f;
      if (L2 && row < row_end) yn = __ldg(p.ynorm + row);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        keyv[r][q] = L2 ? fmaf(2.0f, acc[r][q], -yn) : acc[r][q];
        any = any || (q < nq && row < row_end && keyv[r][q] > thr[q]);
      }
    }
    // steady state: no row of this step beats any threshold -> skip the serial offer loop entirely
    if (!__any_sync(0xffffffffu, any)) continue;
    // offer the RW scores in ascending row order (warp-uniform control flow)
    for (int gg = 0; gg < G; ++gg) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + gg * R + r;
        if (row < row_end) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float key = __shfl_sync(0xffffffffu, keyv[r][q], gg << lpr_log2);
            if (q < nq && key > thr[q]) {
              if (MODE == STREAM_FILTER) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&p.ctl->fcount[q], 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (lane == 0 && slot < STREAM_FCAP) {
                  p.fkey[q * STREAM_FCAP + slot] = key;
                  p.fidx[q * STREAM_FCAP + slot] = row;
                }
              } else {
                top[q].insert(key, row, lane);
                thr[q] = top[q].threshold(kout);
              }
            }
          }
        }
      }
    }
  }
    chunk = int(claim_base + __shfl_sync(0xffffffffu, nxt, 0));
  }

  // ---- LIST: in-block merge, 16 warp lists -> 1 list per query in global memory
  __syncthreads();                                       // queries no longer needed: reuse smem
  STREAM_STAMP(2);
  if (MODE != STREAM_FILTER && KL == 1) {
    // k <= 32: the 16 warp lists fold pairwise (4 levels of merge32) instead of k rounds of a 16-way arg-best
    unsigned long long* sx = reinterpret_cast<unsigned long long*>(sm);     // [STREAM_WARPS][32]
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (q < nq) {                                      // uniform
        const unsigned long long m = block_tree_merge32<STREAM_WARPS>(pack_entry(top[q].key[0], top[q].idx[0]), sx, warp, lane);
        if (warp == 0 && lane < kout) {
          const long long o = ((long long)q * gridDim.x + blockIdx.x) * kout + lane;
          p.cand_key[o] = m ? unordered_f32(uint32_t(m >> 32)) : -CUDART_INF_F;
          p.cand_idx[o] = m ? int(0xFFFFFFFFu - uint32_t(m)) : -1;
        }
        __syncthreads();
      }
    }
  } else if (MODE != STREAM_FILTER) {
    constexpr int LW = 32 * KL;                          // entries per warp list
    float* lk = sm;                                      // [NQ][STREAM_WARPS][LW]
    int* li = reinterpret_cast<int*>(sm + NQ * STREAM_WARPS * LW);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int sl = 0; sl < KL; ++sl) {
        lk[(q * STREAM_WARPS + warp) * LW + lane * KL + sl] = top[q].key[sl];
        li[(q * STREAM_WARPS + warp) * LW + lane * KL + sl] = top[q].idx[sl];
      }
    __syncthreads();
    if (warp < nq) {
      const int q = warp;
      // lane w < STREAM_WARPS owns list w; k rounds of warp arg-best over the heads (key desc, id asc)
      int ptr = 0;
      const bool own = lane < STREAM_WARPS;
      const float* mk = lk + (q * STREAM_WARPS + (own ? lane : 0)) * LW;
      const int* mi = li + (q * STREAM_WARPS + (own ? lane : 0)) * LW;
      for (int r = 0; r < kout; ++r) {
        uint32_t ok = 0; int id = 0x7FFFFFFF; float kv = 0.f;
        if (own && ptr < LW && mi[ptr] >= 0) { kv = mk[ptr]; ok = ordered_f32(kv); id = mi[ptr]; }
        uint32_t wok = ok; int wid = id; int wl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint32_t ook = __shfl_xor_sync(0xffffffffu, wok, o);
          const int oid = __shfl_xor_sync(0xffffffffu, wid, o);
          const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
          if (ook > wok || (ook == wok && oid < wid) || (ook == wok && oid == wid && ol < wl)) { wok = ook; wid = oid; wl = ol; }
        }
        const long long o = ((long long)q * gridDim.x + blockIdx.x) * kout + r;
        if (wok == 0) {
          if (lane == 0) { p.cand_key[o] = -CUDART_INF_F; p.cand_idx[o] = -1; }
        } else if (lane == wl) {
          p.cand_key[o] = kv; p.cand_idx[o] = id;
          ++ptr;
        }
        __syncwarp();
      }
    }
  }

  // ---- the last block to finish produces the final result
  STREAM_STAMP(3);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.ctl->ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  STREAM_STAMP(4);

  if (MODE != STREAM_FILTER && KL == 1) {
    // k <= 32: every warp folds its share of the gridDim.x block lists (all loads issued up-front, one merge32 per
    // list), the 16 partial lists fold through shared memory, warp 0 emits -- ~3 us instead of ~15 us for 148 lists
    unsigned long long* sx = reinterpret_cast<unsigned long long*>(sm);
    const int L = int(gridDim.x);
    for (int q = 0; q < nq; ++q) {
      unsigned long long acc = 0ull;
      for (int l0 = warp; l0 < L; l0 += STREAM_WARPS * 5) {
        float kk[5]; int ii[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int li = l0 + STREAM_WARPS * j;
          const bool ok = li < L && lane < kout;
          const long long o = ((long long)q * L + (ok ? li : 0)) * kout + (ok ? lane : 0);
          kk[j] = __ldcg(p.cand_key + o); ii[j] = ok ? __ldcg(p.cand_idx + o) : -1;
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) acc = merge32(acc, pack_entry(kk[j], ii[j]), lane);
      }
      const unsigned long long m = block_tree_merge32<STREAM_WARPS>(acc, sx, warp, lane);
      if (warp == 0) {
        const float kv = m ? unordered_f32(uint32_t(m >> 32)) : -CUDART_INF_F;
        const int mi = m ? int(0xFFFFFFFFu - uint32_t(m)) : -1;
        if (p.use_pivot_out) {
          // sample pass: only the pivot (key of rank kout - 1, -inf if the sample holds fewer rows)
          if (lane == kout - 1) p.ctl->pivot[q] = kv;
        } else if (lane < kout) {
          const long long o = (long long)q * kout + lane;
          if (p.out_dist) p.out_dist[o] = mi < 0 ? (p.metric_l2 ? CUDART_INF_F : -CUDART_INF_F)
                                                 : (p.metric_l2 ? fmaxf(0.f, s_qnorm[q] - kv) : kv);
          if (p.out_idx) p.out_idx[o] = mi < 0 ? -1ll : (long long)mi + p.id_offset;
          if (p.out_key) p.out_key[o] = kv;
          if (p.out_lbl) p.out_lbl[o] = (mi >= 0 && p.labels) ? __ldg(p.labels + mi) : 0.f;
        }
      }
      __syncthreads();
    }
    if (p.run_if_fallback && threadIdx.x == 0) p.ctl->fallback = 0;
  } else if (MODE != STREAM_FILTER) {
    if (warp < nq) {
      if (p.use_pivot_out) {
        // sample pass: only the pivot (key of rank STREAM_PIVOT_RANK, -inf if the sample holds fewer rows)
        float kth = -CUDART_INF_F;
        merge_lists_warp<int>(p.cand_key, p.cand_idx, nullptr, warp, int(gridDim.x), kout, kout, 0, nullptr, 0, nullptr,
                              nullptr, nullptr, nullptr, nullptr, lane, &kth);
        if (lane == 0) p.ctl->pivot[warp] = kth;
      } else {
        merge_lists_warp<int>(p.cand_key, p.cand_idx, nullptr, warp, int(gridDim.x), kout, kout, p.metric_l2,
                              s_qnorm, p.id_offset, p.labels, p.out_dist, p.out_idx, p.out_lbl, p.out_key, lane, nullptr);
      }
    }
    if (p.run_if_fallback && threadIdx.x == 0) p.ctl->fallback = 0;
  } else {
    // FILTER: sort the appended candidates (key desc, id asc) with a block-wide bitonic sort of packed 64-bit words
    // in shared memory (<= STREAM_FCAP entries = 32 KB) and emit the first kout
    unsigned long long* cw = reinterpret_cast<unsigned long long*>(sm);
    bool failed = false;
    for (int q = 0; q < nq; ++q) {
      const int cnt = __ldcg(&p.ctl->fcount[q]);
      if (cnt > STREAM_FCAP || cnt < min(kout, N)) { failed = true; continue; }
      int np2 = 1;
      while (np2 < cnt) np2 <<= 1;
      __syncthreads();
      for (int i = threadIdx.x; i < np2; i += STREAM_THREADS) {
        unsigned long long wv = 0ull;                       // padding sorts last
        if (i < cnt) {
          const uint32_t ok = ordered_f32(__ldcg(p.fkey + q * STREAM_FCAP + i));
          const uint32_t id = uint32_t(__ldcg(p.fidx + q * STREAM_FCAP + i));
          wv = (static_cast<unsigned long long>(ok) << 32) | (0xFFFFFFFFu - id);   // larger = better (lower id wins ties)
        }
        cw[i] = wv;
      }
      __syncthreads();
      for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = threadIdx.x; i < (np2 >> 1); i += STREAM_THREADS) {
            const int lo = ((i / stride) * (stride << 1)) + (i % stride), hi = lo + stride;
            const bool desc = ((lo & size) == 0);           // descending runs first -> whole array descending
            const unsigned long long a = cw[lo], b = cw[hi];
            if ((a < b) == desc) { cw[lo] = b; cw[hi] = a; }
          }
          __syncthreads();
        }
      }
      for (int r = threadIdx.x; r < kout; r += STREAM_THREADS) {
        const long long o = (long long)q * kout + r;
        if (r < cnt) {
          const unsigned long long wv = cw[r];
          const float kv = unordered_f32(uint32_t(wv >> 32));
          const int mi = int(0xFFFFFFFFu - uint32_t(wv & 0xFFFFFFFFull));
          if (p.out_dist) p.out_dist[o] = p.metric_l2 ? fmaxf(0.f, s_qnorm[q] - kv) : kv;
          p.out_idx[o] = (long long)mi + p.id_offset;
          if (p.out_key) p.out_key[o] = kv;
          if (p.out_lbl) p.out_lbl[o] = p.labels ? p.labels[mi] : 0.f;
        } else {
          // fewer rows than kout in the whole index: pad (faiss convention)
          if (p.out_dist) p.out_dist[o] = p.metric_l2 ? CUDART_INF_F : -CUDART_INF_F;
          p.out_idx[o] = -1;
          if (p.out_key) p.out_key[o] = -CUDART_INF_F;
          if (p.out_lbl) p.out_lbl[o] = 0.f;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) p.ctl->fallback = failed ? 1 : 0;
    if (threadIdx.x < 4) p.ctl->fcount[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) { p.ctl->ticket = 0; p.ctl->next = 0; }
  STREAM_STAMP(5);
  if (p.host_flag && !p.use_pivot_out) {
    // results live in mapped host memory: make every thread's writes visible system-wide, then publish
    // (the block barrier orders every thread's result writes before thread 0's fence: fences are cumulative)
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); *reinterpret_cast<volatile unsigned int*>(p.host_flag) = p.flag_seq; }
  }
}

constexpr size_t stream_smem_bytes(int nq_t, int ld, int mode, int np_leaves, int inline_floats) {
  // prepared queries + leaf sums of the norm + (queries that arrived inside the kernel parameters) their raw copy
  const size_t a = size_t(nq_t) * ld * 4 + size_t(nq_t) * np_leaves * 4 + size_t(inline_floats ? inline_floats + 4 : 0) * 4;
  const size_t b = (mode == STREAM_FILTER) ? size_t(STREAM_FCAP) * 8
                                           : size_t(nq_t) * STREAM_WARPS * 32 * (mode == STREAM_LIST4 ? 4 : 1) * 8;
  return a > b ? a : b;
}

}  // namespace rdb
