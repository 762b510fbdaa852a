// Host-side state shared by the translation units of libradad_flat.so: the index handle, grow-only device scratch,
// error plumbing, and the launchers each kernel family exports to radad_flat.cu (one TU per family so the library
// builds in parallel: the tcgen05, CUDA-core and streaming scorers are ~110 template instantiations between them).
#pragma once
#include "../../include/radad_flat.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>

#include "score_stream.cuh"

namespace rdb {

extern thread_local std::string g_err;     // last error of the calling thread (rdb_last_error(NULL))

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    size_t want = need + need / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); want = need; e = cudaMalloc(&p, want); }
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace rdb

// Per-handle tuning / test options (rdb_set_option); the defaults are the production path.
struct rdb_options {
  int tc_cta_group = 0;          // 0 = auto, 1 | 2 = force cta_group
  int tc_lockstep = 8;           // lock-step window of the TMA producers, in groups of 8 tiles (0 = off)
  int tc_lockstep_spins = 4096;  // polls before a producer gives lock-step up (~5 ms)
  int tc_stages = 64;            // ring slots used (clamped to what the kernel has)
  int tc_query_stationary = 1;   // D <= 256 one-term searches keep the query tile resident
  int tc_pivot = 1;              // sampled admission bound for 32 < k <= 128
  int tc_chunks = 0;             // 0 = cost model, else force the number of database chunks per query tile (A/B)
  int tc_debug = 0;              // RDB_PROFILING builds only: 1 = skip the selection work (results invalid)
  int64_t stream_prof = 0;       // RDB_PROFILING builds only: device address of [blocks][8] u64 phase stamps of the streaming scorer
  int tier1 = 1;                 // fp32 stores: one-term certified pass first
  int tier1_kc = 0;              // 0 = auto, else force 32 | 64 | 128 candidates
  int largek_scorer = 0;         // 0 = auto, 1 = CUDA-core keys, 2 = tensor-core keys
  int64_t largek_rows = 0;       // rows per dense key chunk (0 = default 1M)
  int largek_sample = 1;         // sampled-pivot fast path of the radix select
  int largek_split = 1;          // fp32 stores: split-precision tensor-core keys + certificate for k > 128
  int tc_list10 = 1;             // k <= 10: 10-entry register lists in the tensor-core epilogue (0 = 16 entries)
  int tier1_share2 = 1;          // tier 1 with 32 candidates: two-list cover of 16-entry lists (0 = 32-entry lists)
  int host_pipeline = 1;         // host-buffer searches upload large query batches in pieces behind the running search
};

struct rdb_handle {
  rdb_options opt;
  int d = 0, dp = 0, metric = 0, store = 0, device = 0;
  unsigned flags = 0;
  int64_t n = 0, cap = 0, id_offset = 0, nlabels = 0;
  float* master = nullptr;
  void* hi = nullptr;
  void* lo = nullptr;
  float* ynorm = nullptr;
  float* ynmin32 = nullptr;       // [cap / 32] min |y|^2 over each aligned group of 32 rows (L2 coarse filter of the tcgen05 epilogue)
  float* labels = nullptr;
  void* yext = nullptr;           // [cap][8] bf16: -|y|^2 in three exact bf16 parts (norm slice of the tcgen05 scorer; L2, bf16 operands)
  float cur_hscale = 1.0f;        // scale of the 16-bit query copies of the search in flight (2 = norm-slice scorer)
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // pipelined host-buffer search: second stream for the uploads, two staging buffers, 'uploaded' / 'buffer free' events
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_stage_free[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  int last_algo = 0, last_S = 0;
  bool tc_pivoted = false;        // last tensor-core search used the sampled pivot (needs the completeness check)
  int tc_cg = 1, tc_nqg = 0, tc_S = 0, tc_tpc = 0, tc_ntiles = 0;
  int num_sms = 148;
  int64_t launches = 0;
  int64_t host_syncs = 0;         // times a call blocked the host on the stream (rdb_host_sync_count)
  std::string err;
  std::mutex mu;
  // scratch
  rdb::DevBuf add_stage, q_stage, q_stage2, qf, qhi, qlo, qnorm, cand_key, cand_idx, o_dist, o_idx, o_lbl, ids_stage, rec_stage;
  rdb::DevBuf rr_key, rr_idx, rr_key2, rr_idx2, uncert, fb_qf, fb_qnorm, fb_a, fb_i, fb_l, gthr, tcsync, stream_ctl, fkey, fidx;
  rdb::DevBuf uncert1, t2_qf, t2_qhi, t2_qlo, t2_qnorm, t2_a, t2_i, t2_l;   // fp32 stores: tier-1 list + tier-2 sub-batch
  int t1_level = 0, t1_hold = 0;  // adaptive tier-1 level (exact_split_search) and batches until it decays
  int last_tier1_kc = 0;
  int64_t last_tier1_queries = 0, last_tier1_uncertified = 0;
  // counters of the certified search in flight (read back asynchronously: counts_post / counts_resolve)
  int* pin_counts = nullptr;      // pinned host: [0] tier-1 uncertified, [1] exact-fallback queries
  cudaEvent_t ev_counts = nullptr;
  bool counts_pending = false, pending_tier1 = false;
  int pending_nb = 0, pending_kc1 = 0;
  rdb::DevBuf lk_scores;          // large-k path: dense keys of one (query block x row chunk)
  rdb::DevBuf sat;                // per query: a candidate list was consumed entirely (two-list cover of tier 1)
  rdb::DevBuf qres, res_stage;    // |q - q_hi|^2 per query / per-row residuals of the rows being added (fp32 stores)
  rdb::DevBuf qext;               // [nq][8] bf16 {1, 1, 1, 0, ...}: query side of the norm slice
  int64_t qext_rows = 0;
  rdb::DevBuf dev_ctl;            // device-side control words of the stream-ordered certified search (counts, flags)
  void* pin = nullptr;            // pinned (mapped) host staging of the small-batch path: packed results [+ queries]
  rdb::StreamParams stream_params = {};   // launch parameters of the small-batch search (carry the inline queries)
  unsigned int stream_seq = 0;    // sequence number the latency path publishes in mapped host memory
  size_t pin_bytes = 0;
  float* d_ynorm_max = nullptr;   // [0] max |y|^2 over the shard, [1] max |y - y_hi|^2 (device scalars of the re-rank certificate)
  rdb::DevBuf np_tab;             // piece table of numpy's pairwise summation for rows of d floats (ingest.cuh: NpPlan)
  int np_nleaves = 0, np_nops = 0, np_balanced = 0;
  int64_t last_uncertified = 0;
  bool has_master() const { return store == RDB_STORE_F32 || (flags & RDB_FLAG_KEEP_F32_MASTER); }
  bool has_hi() const { return true; }
  bool has_lo() const { return store == RDB_STORE_F32; }
  bool f16() const { return store == RDB_STORE_F16; }
  // L2 keys straight from the tensor cores (norm slice): bf16 operands only (an f16 part cannot hold |y|^2 > 65504)
  bool use_ext() const { return metric == RDB_METRIC_L2 && store != RDB_STORE_F16; }
};

namespace rdb {

inline int fail(rdb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_err = msg;
  return code;
}
#define CUDA_TRY(h, expr)                                                                              \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      cudaGetLastError();                                                                              \
      return ::rdb::fail(h, e_ == cudaErrorMemoryAllocation ? RDB_ERR_NOMEM : RDB_ERR_CUDA,            \
                         std::string(#expr) + ": " + cudaGetErrorString(e_));                          \
    }                                                                                                  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- launchers exported by the per-family translation units (launch_tc.cu, launch_simt.cu, launch_stream.cu)
struct TcParams;
struct QueryView {
  const float* qf;    // fp32 [nq, D] (fp32 stores)
  const void* qhi;    // 16-bit [nq, Dp]
  const void* qlo;    // 16-bit [nq, Dp] (split-precision)
  const float* qnorm; // [nq]
  int nq;
};

// tcgen05 scorer: encodes the database tensor maps, picks the selector instantiation for k / metric / dump, launches
int launch_tc_cg(rdb_handle* h, TcParams& p, int k, int cg);
int encode_2d(rdb_handle* h, CUtensorMap* m, const void* base, int64_t rows, int D, int Dp, int box_rows);
// exact CUDA-core scorer (selecting form and the k > 128 DUMP form)
int launch_simt(rdb_handle* h, const float* qf, const void* qhi, int nq, int k, int nqt, int S, int rows_per_chunk,
                float* ck, int* ci, const DevPlan* plan = nullptr);
int launch_simt_dump(rdb_handle* h, const QueryView& qv, int q0, int nq, int nqt, int S, int rows_per_chunk, int row0,
                     int row_end, float* dump, long long pitch);
// small-batch streaming scorer
int launch_stream(rdb_handle* h, StreamParams& p, int blocks, int mode);

}  // namespace rdb
