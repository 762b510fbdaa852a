// Large-k selection (128 < k <= 2048 = what faiss-gpu itself accepts): exact selection over a dense key row.
//
// index.search(q, k) -- vector_database.py:181 -- with a k beyond the register / reservoir selectors of the fused
// scorers.  The scorer (tensor cores with the SelectDump epilogue for 16-bit stores, the exact CUDA-core kernel's DUMP
// form for fp32 stores) writes the keys of one block of queries against one chunk of database rows to HBM
// ([queries][rows] fp32, at most ~1 GiB); this kernel then finds, per query, the exact best k of the chunk under the
// library-wide order (key descending, ties by the LOWEST row id) and emits them sorted, as one more candidate list for
// merge_lists_kernel.
//
// Order trick: (key, id) is packed as in pack_cand() -- ordered_f32(key) << 32 | (0xFFFFFFFF - id) -- so all elements
// are DISTINCT 64-bit integers and "the best k" is simply "the k largest".
//
//   exact path   MSB-first radix select (8-bit digits, at most 8 passes over the row, usually 3-4: it stops as soon as
//                the bin holding the k-th element is wanted in full) yields the k-th largest value T; one more pass
//                collects every element >= T (exactly k of them) into shared memory; bitonic sort.
//   fast path    (rows of >= 32768 keys) the same radix select runs on a strided SAMPLE of ~16 K elements and yields a
//                pivot that ~2.5 k (at least 32 sample strides) elements of the row are expected to reach; ONE pass over
//                the row collects everything >= pivot into shared memory (<= 8192 entries); if it found at least k and
//                no more than the buffer holds, sorting them gives the exact answer.  Otherwise (heavy ties, unlucky
//                sample) the exact path runs -- the result never depends on the sample.
#pragma once
#include "common.cuh"

namespace rdb {

constexpr int SELK_THREADS = 1024;
constexpr int SELK_MAXK = 2048;
constexpr int SELK_CAP = 8192;              // shared-memory candidates (64 KB)
constexpr int SELK_SAMPLE = 16384;          // target sample size of the fast path
constexpr int SELK_SAMPLE_MIN_LEN = 32768;
constexpr size_t selk_smem_bytes() { return size_t(SELK_CAP) * 8; }

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

struct SelkShared {
  uint32_t hist[256];
  int bin, remaining, done, cnt;
};

// Returns T with #{ j < n : elem(j) >= T } == want   (1 <= want < n; elements distinct).  elem(j) is evaluated
// once per radix pass.  All threads of the block call it; n_up = n rounded up to a multiple of the block size.
template <class Elem>
__device__ __forceinline__ unsigned long long selk_radix_threshold(Elem elem, int n, int want, SelkShared& sh) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_up = (n + SELK_THREADS - 1) / SELK_THREADS * SELK_THREADS;     // warp-uniform trip counts
  unsigned long long prefix = 0ull, mask = 0ull;
  int remaining = want;
  bool done = false;
  for (int pass = 0; pass < 8 && !done; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) sh.hist[tid] = 0u;
    __syncthreads();
    for (int j = tid; j < n_up; j += SELK_THREADS) {
      const bool in = j < n;
      const unsigned long long v = in ? elem(j) : 0ull;
      const bool match = in && ((v & mask) == prefix);
      selk_hist_add(sh.hist, match, uint32_t(v >> shift) & 255u, lane);
    }
    __syncthreads();
    if (tid == 0) {
      // walk the bins from the top until `remaining` elements are covered
      int cum = 0, b = 255;
      for (; b > 0; --b) {
        if (cum + int(sh.hist[b]) >= remaining) break;
        cum += int(sh.hist[b]);
      }
      sh.bin = b;
      sh.remaining = remaining - cum;
      sh.done = (int(sh.hist[b]) == remaining - cum) ? 1 : 0;   // the whole bin is wanted: no need to look inside it
    }
    __syncthreads();
    prefix |= static_cast<unsigned long long>(sh.bin) << shift;
    mask |= 0xFFull << shift;
    remaining = sh.remaining;
    done = sh.done != 0;
    // (the next pass's hist reset happens only after every thread has read sh.*: it is followed by a barrier, and
    //  thread 0 rewrites sh.* only after two more barriers)
  }
  // v >= prefix  <=>  (v & mask) >= prefix, because prefix is zero outside the mask
  return prefix;
}

// sbuf <- every element of the row that is >= T (at most SELK_CAP are stored); returns how many there are
__device__ __forceinline__ int selk_collect(const float* __restrict__ row, int len, int row0, unsigned long long T,
                                            unsigned long long* sbuf, SelkShared& sh) {
  const int tid = threadIdx.x;
  if (tid == 0) sh.cnt = 0;
  __syncthreads();
  const uint32_t tkey = uint32_t(T >> 32);                 // an element can only reach T if its key part does
  const int len4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (len >> 2) : 0;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  for (int j4 = tid; j4 < len4; j4 += SELK_THREADS) {
    const float4 f = __ldcs(row4 + j4);
    const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (ordered_f32(e[i]) >= tkey) {
        const unsigned long long v = selk_pack(e[i], uint32_t(row0 + 4 * j4 + i));
        if (v >= T) {
          const int slot = atomicAdd(&sh.cnt, 1);
          if (slot < SELK_CAP) sbuf[slot] = v;
        }
      }
    }
  }
  for (int j = 4 * len4 + tid; j < len; j += SELK_THREADS) {
    const unsigned long long v = selk_pack(__ldcs(row + j), uint32_t(row0 + j));
    if (v >= T) {
      const int slot = atomicAdd(&sh.cnt, 1);
      if (slot < SELK_CAP) sbuf[slot] = v;
    }
  }
  __syncthreads();
  return sh.cnt;
}

// scores [nq_blk][pitch]: keys (larger is better) of rows row0 .. row0 + len - 1 of the shard, one row per query of
// this block of queries.  Output: list `chunk` of query (q_first + blockIdx.x) in cand_key / cand_idx laid out
// [q][S][k]: best-first, local row ids, tail slots (when len < k) = (-inf, -1).   Dynamic smem: selk_smem_bytes().
static __global__ void __launch_bounds__(SELK_THREADS) select_dense_kernel(const float* __restrict__ scores, long long pitch,
                                                                    int len, int row0, int k, int S, int chunk,
                                                                    int q_first, int use_sample,
                                                                    float* __restrict__ cand_key,
                                                                    int* __restrict__ cand_idx) {
  extern __shared__ __align__(16) unsigned long long sbuf[];        // [SELK_CAP]
  __shared__ SelkShared sh;
  const int tid = threadIdx.x;
  const float* row = scores + (long long)blockIdx.x * pitch;
  const int kk = min(k, len);

  int cnt = -1;                                                      // candidates in sbuf (>= kk once valid)
  if (use_sample && len >= SELK_SAMPLE_MIN_LEN && len > k) {
    const int st = len / SELK_SAMPLE;                                // >= 2
    const int ns = (len + st - 1) / st;
    const int r = max(32, (5 * kk / 2 + st - 1) / st);               // ~2.5 k row elements expected above the pivot
    if (r < ns) {
      const unsigned long long T = selk_radix_threshold(
          [&](int j) { return selk_pack(__ldg(row + (long long)j * st), uint32_t(row0 + j * st)); }, ns, r, sh);
      const int c = selk_collect(row, len, row0, T, sbuf, sh);
      if (c >= kk && c <= SELK_CAP) cnt = c;                         // sorting these gives the exact best kk
    }
  }
  if (cnt < 0) {
    unsigned long long T = 0ull;                                     // len <= k: everything
    if (len > k)
      T = selk_radix_threshold([&](int j) { return selk_pack(__ldcg(row + j), uint32_t(row0 + j)); }, len, kk, sh);
    cnt = selk_collect(row, len, row0, T, sbuf, sh);                 // exactly kk
  }
  int n2 = 1;
  while (n2 < cnt) n2 <<= 1;
  for (int i = cnt + tid; i < n2; i += SELK_THREADS) sbuf[i] = 0ull;   // below every real element
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
