// Kernel 3 -- small-batch score + select (nq <= 4; the batch-1 streaming latency path, `predict()`:
// pipeline.py:1046-1054, app.py:278).  With one or two queries the search is a pure stream of the database
// through HBM (2*D flops per 2*D or 4*D bytes), so this kernel is built around the memory system instead of the
// tensor cores: every warp reads whole rows with 128-bit coalesced loads (rows are contiguous -> full DRAM
// pages, unlike 128-byte TMA box slices), 4 rows in flight per warp, fp32 FMA against the queries held in
// shared memory, butterfly reduction, and a warp-distributed sorted top-k (lane j holds ranks [KL*j, KL*j+KL); insertion
// is a ballot + shuffle-shift).  The 16 warps of a block merge their lists in shared memory so each block emits
// ONE sorted list per query; merge.cuh folds the gridDim.x lists.  Arithmetic is exact fp32 for fp32 stores
// (no split / re-rank needed) and fp32-accumulated products of the stored 16-bit values otherwise.
#pragma once
#include "common.cuh"
#include "ingest.cuh"

namespace rdb {

constexpr int STREAM_THREADS = 512;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
// rows in flight per warp: 4 fp32 rows or 8 16-bit rows (~12 KB of loads outstanding per warp either way at D = 768)

template <typename T> struct StreamVec;
template <> struct StreamVec<float> {
  static constexpr int EPV = 4;   // elements per 128-bit load
  static constexpr int R = 4;     // rows in flight per warp
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
};
template <> struct StreamVec<__nv_bfloat16> {
  static constexpr int EPV = 8;
  static constexpr int R = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};
template <> struct StreamVec<__half> {
  static constexpr int EPV = 8;
  static constexpr int R = 8;
  __device__ static __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
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

// Y [N, ld] stored rows (T = fp32 master or the 16-bit store), ld multiple of EPV, columns [D, ld) zero (16-bit) --
// D itself must be a multiple of EPV for fp32 (checked by the host; otherwise another scorer is used).
// Qs: queries as fp32 [NQ][ld] in global memory (already normalised / rounded to the store dtype).
// cand_* [nq][gridDim.x][kout].
template <typename T, int NQ, bool L2, int KL>
__global__ void __launch_bounds__(STREAM_THREADS, 1) score_select_stream_kernel(
    const T* __restrict__ Y, int ld, const float* __restrict__ ynorm, int N, const float* __restrict__ Qs, int nq,
    int rows_per_block, float* __restrict__ cand_key, int* __restrict__ cand_idx, int kout, int lpr_log2) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                                        // [NQ][ld]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < NQ * ld; i += STREAM_THREADS) qs[i] = (i / ld < nq) ? Qs[i] : 0.f;
  __syncthreads();

  constexpr int EPV = StreamVec<T>::EPV;
  const int nvec = ld / EPV;                             // 128-bit vectors per row
  const int row_begin = blockIdx.x * rows_per_block;
  const int row_end = min(N, row_begin + rows_per_block);

  WarpList<KL> top[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) top[q].init();
  float thr[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) thr[q] = -CUDART_INF_F;

  // Lane groups of `lpr` lanes own rows (short rows: 8 or 16 lanes per row, so the shuffle reduction is amortised over
  // several rows per warp step).  Per step a warp takes RW = R * G consecutive rows: group g rows r0 + g*R .. + R.
  constexpr int R = StreamVec<T>::R;
  const int lpr = 1 << lpr_log2, G = 32 >> lpr_log2;
  const int grp = lane >> lpr_log2, l = lane & (lpr - 1);
  const int RW = R * G;
  for (int r0 = row_begin + warp * RW; r0 < row_end; r0 += STREAM_WARPS * RW) {
    float acc[R][NQ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[r][q] = 0.f;
    for (int c = l; c < nvec; c += lpr) {
      float y[R][EPV];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(r0 + grp * R + r, row_end - 1);   // clamp: duplicates are discarded below
        StreamVec<T>::load(Y + (long long)row * ld + c * EPV, y[r]);
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float qv[EPV];
#pragma unroll
        for (int e = 0; e < EPV; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(qs + q * ld + c * EPV + e);
          qv[e] = t.x; qv[e + 1] = t.y; qv[e + 2] = t.z; qv[e + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < EPV; ++e) acc[r][q] = fmaf(qv[e], y[r][e], acc[r][q]);
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[r][q] += __shfl_xor_sync(0xffffffffu, acc[r][q], o);
    // offer the RW scores in ascending row order (warp-uniform control flow)
    for (int gg = 0; gg < G; ++gg) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + gg * R + r;
        if (row < row_end) {
          float yn = 0.f;
          if (L2) yn = __ldg(ynorm + row);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float sc = __shfl_sync(0xffffffffu, acc[r][q], gg << lpr_log2);
            const float key = L2 ? fmaf(2.0f, sc, -yn) : sc;
            if (q < nq && key > thr[q]) {
              top[q].insert(key, row, lane);
              thr[q] = top[q].threshold(kout);
            }
          }
        }
      }
    }
  }

  // ---- in-block merge: 16 warp lists -> 1 list per query
  __syncthreads();                                       // queries no longer needed: reuse smem
  constexpr int LW = 32 * KL;                            // entries per warp list
  float* lk = sm;                                        // [NQ][STREAM_WARPS][LW]
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
        if (lane == 0) { cand_key[o] = -CUDART_INF_F; cand_idx[o] = -1; }
      } else if (lane == wl) {
        cand_key[o] = kv; cand_idx[o] = id;
        ++ptr;
      }
      __syncwarp();
    }
  }
}

constexpr size_t stream_smem_bytes(int nq_t, int ld, int kl) {
  const size_t a = size_t(nq_t) * ld * 4, b = size_t(nq_t) * STREAM_WARPS * 32 * kl * 8;
  return a > b ? a : b;
}

}  // namespace rdb
