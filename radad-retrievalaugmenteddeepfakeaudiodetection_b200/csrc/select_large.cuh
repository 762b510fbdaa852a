// Large-k selection (128 < k <= 2048 = what faiss-gpu itself accepts): exact radix select over a dense key row.
//
// index.search(q, k) -- vector_database.py:181 -- with a k beyond the register / reservoir selectors of the fused
// scorers.  The exact CUDA-core scorer (score_simt.cuh, DUMP form) writes the keys of one block of queries against one
// chunk of database rows to HBM ([queries][rows] fp32, at most ~1 GiB); this kernel then finds, per query, the exact
// best k of the chunk under the library-wide order (key descending, ties by the LOWEST row id) and emits them sorted,
// as one more candidate list for merge_lists_kernel.  HBM-bound: the chunk's keys are read once per radix pass.
//
// Order trick: (key, id) is packed as in pack_cand() -- ordered_f32(key) << 32 | (0xFFFFFFFF - id) -- so all elements
// are DISTINCT 64-bit integers and "the best k" is simply "the k largest": an MSB-first radix select (8-bit digits,
// at most 8 passes, usually 3-4: it stops as soon as the bin holding the k-th element is needed in full) yields the
// k-th largest value T, one more pass collects every element >= T (exactly k of them) into shared memory, and a
// bitonic sort orders them.
#pragma once
#include "common.cuh"

namespace rdb {

constexpr int SELK_THREADS = 1024;
constexpr int SELK_MAXK = 2048;

__device__ __forceinline__ unsigned long long selk_pack(float key, uint32_t row) {
  return (static_cast<unsigned long long>(ordered_f32(key)) << 32) |
         static_cast<unsigned long long>(0xFFFFFFFFu - row);
}

// histogram update for one element per lane; the digit most lanes share (pass 0: sign + exponent bits) is counted
// with one atomic per warp, the rest with plain shared-memory atomics
__device__ __forceinline__ void selk_hist_add(uint32_t* hist, bool valid, uint32_t digit, int lane) {
  unsigned act = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    if (act == 0u) return;                       // warp-uniform
    const int leader = __ffs(act) - 1;
    const uint32_t d0 = __shfl_sync(0xffffffffu, digit, leader);
    const unsigned same = __ballot_sync(0xffffffffu, valid && digit == d0);
    if (lane == leader) atomicAdd(&hist[d0], uint32_t(__popc(same)));
    if (valid && digit == d0) valid = false;
    act &= ~same;
  }
  if (valid) atomicAdd(&hist[digit], 1u);
}

// scores [nq_blk][pitch]: keys (larger is better) of rows row0 .. row0 + len - 1 of the shard, one row per query of
// this block of queries.  Output: list `chunk` of query (q_first + blockIdx.x) in cand_key / cand_idx laid out
// [q][S][k]: best-first, local row ids, tail slots (when len < k) = (-inf, -1).
__global__ void __launch_bounds__(SELK_THREADS) select_dense_kernel(const float* __restrict__ scores, long long pitch,
                                                                    int len, int row0, int k, int S, int chunk,
                                                                    int q_first, float* __restrict__ cand_key,
                                                                    int* __restrict__ cand_idx) {
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sbuf[SELK_MAXK];
  __shared__ int s_bin, s_remaining, s_done, s_cnt;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* row = scores + (long long)blockIdx.x * pitch;
  const int kk = min(k, len);
  const int len_up = (len + SELK_THREADS - 1) / SELK_THREADS * SELK_THREADS;   // warp-uniform trip counts

  unsigned long long prefix = 0ull, mask = 0ull;
  if (len > k) {
    int remaining = kk;
    bool done = false;
    for (int pass = 0; pass < 8 && !done; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) hist[tid] = 0u;
      __syncthreads();
      for (int j = tid; j < len_up; j += SELK_THREADS) {
        const bool in = j < len;
        const unsigned long long v = in ? selk_pack(__ldcg(row + j), uint32_t(row0 + j)) : 0ull;
        const bool match = in && ((v & mask) == prefix);
        selk_hist_add(hist, match, uint32_t(v >> shift) & 255u, lane);
      }
      __syncthreads();
      if (tid == 0) {
        // walk the bins from the top until `remaining` elements are covered
        int cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + int(hist[b]) >= remaining) break;
          cum += int(hist[b]);
        }
        s_bin = b;
        s_remaining = remaining - cum;
        s_done = (int(hist[b]) == remaining - cum) ? 1 : 0;   // the whole bin is wanted: no need to look inside it
      }
      __syncthreads();
      prefix |= static_cast<unsigned long long>(s_bin) << shift;
      mask |= 0xFFull << shift;
      remaining = s_remaining;
      done = s_done != 0;
      // (the next pass's hist reset happens only after every thread has read s_*: it is followed by a barrier, and
      //  thread 0 rewrites s_* only after two more barriers)
    }
  }
  // collect: v >= prefix  <=>  (v & mask) >= prefix, because prefix is zero outside the mask; exactly kk elements
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int j = tid; j < len; j += SELK_THREADS) {
    const unsigned long long v = selk_pack(__ldcg(row + j), uint32_t(row0 + j));
    if (v >= prefix) {
      const int slot = atomicAdd(&s_cnt, 1);
      if (slot < SELK_MAXK) sbuf[slot] = v;
    }
  }
  __syncthreads();
  int n2 = 1;
  while (n2 < kk) n2 <<= 1;
  for (int i = kk + tid; i < n2; i += SELK_THREADS) sbuf[i] = 0ull;     // below every real element
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (n2 >> 1); i += SELK_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sbuf[lo], b = sbuf[hi];
        if ((a < b) == desc) { sbuf[lo] = b; sbuf[hi] = a; }
      }
      __syncthreads();
    }
  }
  const long long base = ((long long)(q_first + blockIdx.x) * S + chunk) * (long long)k;
  for (int j = tid; j < k; j += SELK_THREADS) {
    if (j < kk) {
      const unsigned long long v = sbuf[j];
      cand_key[base + j] = unordered_f32(uint32_t(v >> 32));
      cand_idx[base + j] = int(0xFFFFFFFFu - uint32_t(v));
    } else {
      cand_key[base + j] = -CUDART_INF_F;
      cand_idx[base + j] = -1;
    }
  }
}

}  // namespace rdb
