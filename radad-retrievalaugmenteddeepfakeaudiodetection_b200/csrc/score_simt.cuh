// Exact-fp32 score + select on CUDA cores (kernel "v0" / exact path).
// Computes what faiss IndexFlat{L2,IP}.search computes (vector_database.py:181) for fp32 storage with fp32
// FMA accumulation, never materialising the [nq, N] distance matrix: each block owns a 128-query tile and a
// contiguous chunk of database rows, and every thread keeps a register-resident running top-KT list.
//
// Used for: fp32 stores (parity tolerance 1e-5), tiny databases, odd D, and as the cross-check for the
// tcgen05 kernel.  Not the headline path (that is score_tc.cuh).
#pragma once
#include "common.cuh"

namespace rdb {

constexpr int SIMT_BM = 128;   // queries per block tile
constexpr int SIMT_BN = 128;   // database rows per tile
constexpr int SIMT_BK = 16;
constexpr int SIMT_LD = 132;   // padded leading dimension of the transposed operand tiles
constexpr int SIMT_SLD = 129;  // padded leading dimension of the score tile
constexpr int SIMT_LISTS = 2;  // candidate lists produced per (query, chunk): two column halves

constexpr size_t simt_smem_bytes() {
  return sizeof(float) * (size_t(2) * 2 * SIMT_BK * SIMT_LD + size_t(SIMT_BM) * SIMT_SLD + SIMT_BN);
}

// Loads 4 consecutive K elements of row `row` (pitch ld) as fp32; out-of-range rows / columns read as 0.
template <typename T, bool ALIGNED>
__device__ __forceinline__ float4 simt_load4(const T* __restrict__ base, long long row, long long nrows, int k, int D,
                                             int ld);

template <>
__device__ __forceinline__ float4 simt_load4<float, true>(const float* __restrict__ base, long long row,
                                                          long long nrows, int k, int D, int ld) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows && k < D) v = __ldg(reinterpret_cast<const float4*>(base + row * (long long)ld + k));
  return v;
}
template <>
__device__ __forceinline__ float4 simt_load4<float, false>(const float* __restrict__ base, long long row,
                                                           long long nrows, int k, int D, int ld) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows) {
    const float* p = base + row * (long long)ld + k;
    if (k + 0 < D) v.x = __ldg(p + 0);
    if (k + 1 < D) v.y = __ldg(p + 1);
    if (k + 2 < D) v.z = __ldg(p + 2);
    if (k + 3 < D) v.w = __ldg(p + 3);
  }
  return v;
}
// 16-bit stores: pitch ld is a multiple of 8 and columns [D, ld) are zero, so 8-byte loads are always legal.
template <>
__device__ __forceinline__ float4 simt_load4<__nv_bfloat16, true>(const __nv_bfloat16* __restrict__ base,
                                                                  long long row, long long nrows, int k, int D,
                                                                  int ld) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows && k < ld) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(base + row * (long long)ld + k));
    v.x = __uint_as_float(u.x << 16); v.y = __uint_as_float(u.x & 0xFFFF0000u);
    v.z = __uint_as_float(u.y << 16); v.w = __uint_as_float(u.y & 0xFFFF0000u);
  }
  return v;
}
template <>
__device__ __forceinline__ float4 simt_load4<__half, true>(const __half* __restrict__ base, long long row,
                                                           long long nrows, int k, int D, int ld) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows && k < ld) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(base + row * (long long)ld + k));
    const __half2 a = *reinterpret_cast<const __half2*>(&u.x), b = *reinterpret_cast<const __half2*>(&u.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    v = make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  return v;
}

// Q [nq, D] and Y [N, D] of element type T (fp32, or the 16-bit store type), row pitch ld, K contiguous.
// ynorm[N] = |y|^2 of the stored values (L2 only).
// cand_key / cand_idx : [nq][S * SIMT_LISTS][kout]   (key: larger is better; idx: local row id or -1)
// grid.x = nqt * S, block = 256.   chunk c covers rows [c*rows_per_chunk, min(N, (c+1)*rows_per_chunk)).
//
// DUMP form (k > 128, see select_dense_kernel in select_large.cuh): no selection; the keys of rows
// [dump_row0, N) are written to dump[q][row - dump_row0] (row pitch dump_pitch floats) with exactly the arithmetic of
// the selecting form, chunk c covering rows [dump_row0 + c*rows_per_chunk, ...).
template <int KT, bool L2, typename T, bool ALIGNED, bool DUMP = false>
__global__ void __launch_bounds__(256) score_select_simt_kernel(const T* __restrict__ Q,
                                                                const T* __restrict__ Y,
                                                                const float* __restrict__ ynorm, int nq, int N, int D,
                                                                int ld,
                                                                int nqt, int S, int rows_per_chunk,
                                                                float* __restrict__ cand_key,
                                                                int* __restrict__ cand_idx, int kout,
                                                                float* __restrict__ dump = nullptr,
                                                                long long dump_pitch = 0, int dump_row0 = 0,
                                                                const DevPlan* __restrict__ plan = nullptr) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                                   // [2][BK][LD]
  float* Bs = As + 2 * SIMT_BK * SIMT_LD;             // [2][BK][LD]
  float* Ss = Bs + 2 * SIMT_BK * SIMT_LD;             // [BM][SLD]
  float* Yn = Ss + SIMT_BM * SIMT_SLD;                // [BN]

  const int tid = threadIdx.x;
  // device-sized launch (DevPlan): the grid covers the worst case, the real extent comes from the plan
  int num_units = nqt * S;
  if (plan) {
    nq = __ldcg(&plan->nq); nqt = __ldcg(&plan->s_nqt); S = __ldcg(&plan->s_S); rows_per_chunk = __ldcg(&plan->s_rows);
    num_units = __ldcg(&plan->s_units);
    if (nq <= 0) return;
  }
  const int ty = tid >> 4, tx = tid & 15;
  // loader mapping: two float4 per operand per thread
  const int lrow0 = tid >> 2, lk = (tid & 3) * 4;     // rows lrow0 and lrow0 + 64
  // scan mapping: thread -> (query row, column half)
  const int srow = tid >> 1, shalf = tid & 1;
  const int nk = (D + SIMT_BK - 1) / SIMT_BK;
  for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
  const int qtile = unit % nqt;
  const int chunk = unit / nqt;
  const long long q0 = (long long)qtile * SIMT_BM;
  const int row_begin = (DUMP ? dump_row0 : 0) + chunk * rows_per_chunk;
  const int row_end = min(N, row_begin + rows_per_chunk);
  TopK<KT> top;
  top.init();

  for (int n0 = row_begin; n0 < row_end; n0 += SIMT_BN) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    ra[0] = simt_load4<T, ALIGNED>(Q, q0 + lrow0, nq, lk, D, ld);
    ra[1] = simt_load4<T, ALIGNED>(Q, q0 + lrow0 + 64, nq, lk, D, ld);
    rb[0] = simt_load4<T, ALIGNED>(Y, (long long)n0 + lrow0, row_end, lk, D, ld);
    rb[1] = simt_load4<T, ALIGNED>(Y, (long long)n0 + lrow0 + 64, row_end, lk, D, ld);

    for (int kb = 0; kb < nk; ++kb) {
      float* a_s = As + (kb & 1) * SIMT_BK * SIMT_LD;
      float* b_s = Bs + (kb & 1) * SIMT_BK * SIMT_LD;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lrow0 + 64 * h;
        a_s[(lk + 0) * SIMT_LD + r] = ra[h].x; a_s[(lk + 1) * SIMT_LD + r] = ra[h].y;
        a_s[(lk + 2) * SIMT_LD + r] = ra[h].z; a_s[(lk + 3) * SIMT_LD + r] = ra[h].w;
        b_s[(lk + 0) * SIMT_LD + r] = rb[h].x; b_s[(lk + 1) * SIMT_LD + r] = rb[h].y;
        b_s[(lk + 2) * SIMT_LD + r] = rb[h].z; b_s[(lk + 3) * SIMT_LD + r] = rb[h].w;
      }
      __syncthreads();
      if (kb + 1 < nk) {
        const int k1 = (kb + 1) * SIMT_BK + lk;
        ra[0] = simt_load4<T, ALIGNED>(Q, q0 + lrow0, nq, k1, D, ld);
        ra[1] = simt_load4<T, ALIGNED>(Q, q0 + lrow0 + 64, nq, k1, D, ld);
        rb[0] = simt_load4<T, ALIGNED>(Y, (long long)n0 + lrow0, row_end, k1, D, ld);
        rb[1] = simt_load4<T, ALIGNED>(Y, (long long)n0 + lrow0 + 64, row_end, k1, D, ld);
      }
#pragma unroll
      for (int k = 0; k < SIMT_BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(a_s + k * SIMT_LD + ty * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(a_s + k * SIMT_LD + 64 + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4*>(b_s + k * SIMT_LD + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(b_s + k * SIMT_LD + 64 + tx * 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      // double-buffered smem: the next iteration writes the other buffer, one barrier per step suffices
    }
    __syncthreads();  // previous tile's scan must be done before Ss / Yn are overwritten (also covers last k-step)
    if (L2 && tid < SIMT_BN) Yn[tid] = (n0 + tid < row_end) ? __ldg(ynorm + n0 + tid) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty * 4 + (i & 3) + 64 * (i >> 2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = tx * 4 + (j & 3) + 64 * (j >> 2);
        Ss[r * SIMT_SLD + c] = acc[i][j];
      }
    }
    __syncthreads();
    if (DUMP) {
      // one warp per query row of the tile, 4 coalesced 128-byte stores per row
      const int w = tid >> 5, lane = tid & 31;
      for (int rr = w; rr < SIMT_BM && q0 + rr < nq; rr += 8) {
        float* drow = dump + (q0 + rr) * dump_pitch + (n0 - dump_row0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = lane + 32 * i;
          float v = Ss[rr * SIMT_SLD + c];
          if (L2) v = fmaf(2.0f, v, -Yn[c]);
          if (n0 + c < row_end) drow[c] = v;
        }
      }
    } else {
      const float* srow_p = Ss + srow * SIMT_SLD + shalf * 64;
      const int cbase = n0 + shalf * 64;
#pragma unroll 4
      for (int j = 0; j < 64; ++j) {
        float v = srow_p[j];
        if (L2) v = fmaf(2.0f, v, -Yn[shalf * 64 + j]);
        if (cbase + j < row_end && v > top.worst()) top.insert(v, cbase + j);
      }
    }
    // Ss / Yn are rewritten only after the next tile's k-loop barriers, As/Bs after the barrier above.
  }

  const long long q = q0 + srow;
  if (!DUMP && q < nq) {
    const long long base = ((q * S + chunk) * SIMT_LISTS + shalf) * (long long)kout;
#pragma unroll
    for (int j = 0; j < KT; ++j)
      if (j < kout) { cand_key[base + j] = top.key[j]; cand_idx[base + j] = top.idx[j]; }
  }
  __syncthreads();      // the next unit of this block reuses the shared tiles
  }
}

}  // namespace rdb
