// Host side of the C ABI declared in include/radad_flat.h (sm_100a only; no CPU fallback).
//
// Device layout of one index (one GPU, one row shard):
//   master f32 [cap, D]     fp32 rows            (RDB_STORE_F32, or 16-bit stores with RDB_FLAG_KEEP_F32_MASTER)
//   hi     16b [cap, Dp]    bf16/f16 rows, K-major, pitch Dp = round_up(D, 8) (TMA needs 16-byte row strides)
//   lo     16b [cap, Dp]    bf16(x - hi) residual (RDB_STORE_F32 only: split-precision tensor-core scorer)
//   ynorm  f32 [cap]        |y|^2 of the values the scorer sees;   cap is a multiple of 256 (tile padding)
//   labels f32 [n]          neighbour labels for the kNN vote
// Everything a search needs beyond that lives in grow-only scratch owned by the handle, so a steady-state
// search performs no allocation.
#include "handle.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "ingest.cuh"
#include "merge.cuh"
#include "score_simt.cuh"
#include "score_stream.cuh"
#include "score_tc.cuh"
#include "select_large.cuh"

using namespace rdb;

namespace rdb { thread_local std::string g_err; }

namespace {

// every place a call blocks the host on the handle's stream goes through here (rdb_host_sync_count)
cudaError_t host_sync(rdb_handle* h) { h->host_syncs++; return cudaStreamSynchronize(h->stream); }

// ---------------------------------------------------------------------------------------------- storage
int grow_to(rdb_handle* h, int64_t need, bool exact = false) {
  if (need <= h->cap) return RDB_OK;
  int64_t cap = exact ? need : std::max<int64_t>(need, h->cap + h->cap / 2);
  cap = round_up(std::max<int64_t>(cap, 256), 256);
  if (cap >= (int64_t(1) << 31)) return fail(h, RDB_ERR_UNSUPPORTED, "more than 2^31-1 rows per shard");
  const size_t D = h->d, Dp = h->dp;
  float* master = nullptr; void* hi = nullptr; void* lo = nullptr; float* ynorm = nullptr; float* ynmin32 = nullptr;
  void* yext = nullptr;
  auto cleanup = [&] { cudaFree(master); cudaFree(hi); cudaFree(lo); cudaFree(ynorm); cudaFree(ynmin32); cudaFree(yext); cudaGetLastError(); };
  cudaError_t e = cudaSuccess;
  if (h->has_master()) e = cudaMalloc(&master, size_t(cap) * D * 4);
  if (e == cudaSuccess) e = cudaMalloc(&hi, size_t(cap) * Dp * 2);
  if (e == cudaSuccess && h->has_lo()) e = cudaMalloc(&lo, size_t(cap) * Dp * 2);
  if (e == cudaSuccess) e = cudaMalloc(&ynorm, size_t(cap) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&ynmin32, size_t(cap / 32) * 4);
  if (e == cudaSuccess && h->use_ext()) e = cudaMalloc(&yext, size_t(cap) * 16);
  if (e != cudaSuccess) {
    cleanup();
    return fail(h, RDB_ERR_NOMEM, std::string("device allocation for ") + std::to_string(cap) + " rows failed: " +
                                      cudaGetErrorString(e));
  }
  cudaStream_t s = h->stream;
  if (h->n > 0) {
    if (master) cudaMemcpyAsync(master, h->master, size_t(h->n) * D * 4, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(hi, h->hi, size_t(h->n) * Dp * 2, cudaMemcpyDeviceToDevice, s);
    if (lo) cudaMemcpyAsync(lo, h->lo, size_t(h->n) * Dp * 2, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(ynorm, h->ynorm, size_t(h->n) * 4, cudaMemcpyDeviceToDevice, s);
  }
  cudaMemsetAsync(ynorm + h->n, 0, size_t(cap - h->n) * 4, s);
  cudaMemsetAsync(ynmin32, 0, size_t(cap / 32) * 4, s);
  if (h->n > 0) cudaMemcpyAsync(ynmin32, h->ynmin32, size_t((h->n + 31) / 32) * 4, cudaMemcpyDeviceToDevice, s);
  if (h->n > 0 && yext) cudaMemcpyAsync(yext, h->yext, size_t(h->n) * 16, cudaMemcpyDeviceToDevice, s);
  cudaError_t es = cudaStreamSynchronize(s);
  if (es != cudaSuccess) { cleanup(); return fail(h, RDB_ERR_CUDA, std::string("grow copy: ") + cudaGetErrorString(es)); }
  cudaFree(h->master); cudaFree(h->hi); cudaFree(h->lo); cudaFree(h->ynorm); cudaFree(h->ynmin32); cudaFree(h->yext);
  h->master = master; h->hi = hi; h->lo = lo; h->ynorm = ynorm; h->ynmin32 = ynmin32; h->yext = yext; h->cap = cap;
  return RDB_OK;
}

// numpy's pairwise_sum recursion for a row of n floats (numpy/_core/src/umath/loops_utils.h.src: PW_BLOCKSIZE = 128,
// left part = n / 2 rounded down to a multiple of 8): leaves in left-to-right order + the post-order fold program.
int np_plan_rec(int off, int n, std::vector<int>& offs, std::vector<int>& lens, std::vector<int>& ops, bool* balanced) {
  if (n <= 128) { offs.push_back(off); lens.push_back(n); return int(offs.size()) - 1; }
  int n2 = n / 2;
  n2 -= n2 % 8;
  const int a = np_plan_rec(off, n2, offs, lens, ops, balanced);
  const int b = np_plan_rec(off + n2, n - n2, offs, lens, ops, balanced);
  if (b - a != int(offs.size()) - b) *balanced = false;      // left and right subtree hold different numbers of leaves
  ops.push_back(a); ops.push_back(b);
  return a;
}
int np_plan_build(rdb_handle* h) {
  std::vector<int> offs, lens, ops;
  bool balanced = true;
  np_plan_rec(0, h->d, offs, lens, ops, &balanced);
  for (size_t i = 1; i < lens.size(); ++i) balanced = balanced && lens[i] == lens[0];      // ... of EQUAL pieces
  h->np_balanced = balanced ? 1 : 0;
  std::vector<int> tab(offs);
  tab.insert(tab.end(), lens.begin(), lens.end());
  tab.insert(tab.end(), ops.begin(), ops.end());
  CUDA_TRY(h, h->np_tab.ensure(tab.size() * 4));
  CUDA_TRY(h, cudaMemcpy(h->np_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  h->np_nleaves = int(offs.size()); h->np_nops = int(ops.size() / 2);
  return RDB_OK;
}
NpPlan np_plan(rdb_handle* h) { return NpPlan{h->np_tab.as<int>(), h->np_nleaves, h->np_nops, h->np_balanced, h->d}; }

// launch the fused ingest kernel (also used to prepare queries)
int launch_ingest(rdb_handle* h, const float* x, int64_t n, int normalize, int norm_of_hi, float* master, void* hi,
                  void* lo, float* norm2, float hscale = 1.0f, float* res2 = nullptr, const int* n_dev = nullptr) {
  if (n <= 0) return RDB_OK;
  const int D = h->d, Dp = h->dp;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  // one warp per row; few rows (a batch of 256 queries of 5376 floats is 32 blocks of 8 warps: 28 us on 32 SMs) are
  // spread over the SMs with fewer warps per block
  const int warps_per_block = n >= int64_t(h->num_sms) * 8 ? 8 : int(std::max<int64_t>(1, (n + h->num_sms - 1) / h->num_sms));
  int64_t blocks = std::min<int64_t>((n + warps_per_block - 1) / warps_per_block, int64_t(h->num_sms) * 16);
  dim3 grid((unsigned)blocks), block(32 * warps_per_block);
  cudaStream_t s = h->stream;
  // rows of 128 * NC floats with a store layout the specialised kernel knows: hi only / master + hi + lo / master + hi
  {
    const int mode = (!master && hi && !lo && norm_of_hi) ? 0 : ((master && hi && lo && !norm_of_hi) ? 1
                     : ((master && hi && !lo && norm_of_hi) ? 2 : -1));
    if (vec4 && mode >= 0 && norm2 && D == Dp && D % 128 == 0 && D <= 1024) {
      const int ncf = D / 128;
#define FAST_LAUNCH(T16, NC, NORM, MODE)                                                                                  \
      ingest_fast_kernel<T16, NC, NORM, MODE><<<grid, block, NORM ? warps_per_block * FastShape<NC>::WARP_FLOATS * 4 : 0, \
                                                h->stream>>>(x, n, master, (T16*)hi, (T16*)lo, norm2, hscale, res2, n_dev)
#define FAST_MODE(T16, NC, NORM)                                                                                          \
      do { if (mode == 0) FAST_LAUNCH(T16, NC, NORM, 0); else if (mode == 1) FAST_LAUNCH(T16, NC, NORM, 1);               \
           else FAST_LAUNCH(T16, NC, NORM, 2); } while (0)
#define FAST_NORM(T16, NC) do { if (normalize) FAST_MODE(T16, NC, true); else FAST_MODE(T16, NC, false); } while (0)
#define FAST_NC(T16)                                                                                                      \
      do { switch (ncf) { case 1: FAST_NORM(T16, 1); break; case 2: FAST_NORM(T16, 2); break; case 3: FAST_NORM(T16, 3); break; \
                          case 4: FAST_NORM(T16, 4); break; case 5: FAST_NORM(T16, 5); break; case 6: FAST_NORM(T16, 6); break; \
                          case 7: FAST_NORM(T16, 7); break; default: FAST_NORM(T16, 8); break; } } while (0)
      if (h->f16()) FAST_NC(__half); else FAST_NC(__nv_bfloat16);
#undef FAST_NC
#undef FAST_NORM
#undef FAST_MODE
#undef FAST_LAUNCH
      h->launches++;
      CUDA_TRY(h, cudaGetLastError());
      return RDB_OK;
    }
  }
  // rows of up to 1024 floats are held in registers (one read, NC independent 128-bit loads per lane)
  int nc = !vec4 ? 0 : (Dp <= 256 ? 2 : (Dp <= 512 ? 4 : (Dp <= 1024 ? 8 : 0)));
  if (normalize && h->np_nleaves > NP_MAX_REG_LEAVES) nc = 0;
  // normalising launches stage the row's squares / leaf sums of numpy's summation order in shared memory (ingest.cuh)
  const size_t smem = normalize ? size_t(warps_per_block) * ingest_warp_floats(h->np_nleaves, D, nc > 0) * 4 : 0;
  if (smem > 200 * 1024) return fail(h, RDB_ERR_UNSUPPORTED, "normalisation of rows this long is not supported");
  const NpPlan np = np_plan(h);
#define INGEST_LAUNCH(T16, V4, NC, NORM)                                                                                  \
  do {                                                                                                                    \
    if (smem > 48 * 1024)                                                                                                 \
      CUDA_TRY(h, cudaFuncSetAttribute(ingest_rows_kernel<T16, V4, NC, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       int(smem)));                                                                       \
    ingest_rows_kernel<T16, V4, NC, NORM><<<grid, block, smem, s>>>(x, n, D, Dp, norm_of_hi, master, (T16*)hi, (T16*)lo,  \
                                                                    norm2, np, hscale, res2, n_dev);                      \
  } while (0)
#define INGEST_N(T16, V4, NC)                                                                             \
  do { if (normalize) INGEST_LAUNCH(T16, V4, NC, true); else INGEST_LAUNCH(T16, V4, NC, false); } while (0)
#define INGEST_T(T16)                                                                                     \
  do {                                                                                                    \
    if (nc == 2) INGEST_N(T16, true, 2); else if (nc == 4) INGEST_N(T16, true, 4);                        \
    else if (nc == 8) INGEST_N(T16, true, 8); else if (vec4) INGEST_N(T16, true, 0);                      \
    else INGEST_N(T16, false, 0);                                                                         \
  } while (0)
  if (h->f16()) INGEST_T(__half); else INGEST_T(__nv_bfloat16);
#undef INGEST_N
#undef INGEST_T
#undef INGEST_LAUNCH
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

// ---------------------------------------------------------------------------------------------- scheduling
// Split the database into S chunks so that nqt * S work units fill the machine in whole waves.
// cost model: waves * (tiles_per_chunk + per-unit overhead);  `slots` = concurrently resident CTAs.
int choose_splits(int64_t nqt, int64_t ntiles, int slots, int max_lists, int min_tiles, int* tiles_per_chunk,
                  double unit_overhead_tiles = 2.0) {
  int64_t maxS = std::min<int64_t>(std::min<int64_t>(max_lists, ntiles), std::max<int64_t>(1, ntiles / min_tiles));
  double best = 1e300;
  int bestS = 1;
  int64_t best_tpc = ntiles;
  for (int64_t S = 1; S <= maxS; ++S) {
    const int64_t tpc = (ntiles + S - 1) / S;
    const int64_t S2 = (ntiles + tpc - 1) / tpc;
    if (S2 != S) continue;
    const int64_t units = nqt * S2;
    const int64_t waves = (units + slots - 1) / slots;
    const double cost = double(waves) * (double(tpc) + unit_overhead_tiles);
    if (cost < best * 0.999) { best = cost; bestS = int(S2); best_tpc = tpc; }
  }
  *tiles_per_chunk = int(best_tpc);
  return bestS;
}

// CTAs per MMA group.  Both forms are built and parity-tested: cta_group::1 (one CTA = one 128-query tile, M = 128)
// and cta_group::2 (a CTA pair runs M = 256 MMAs, each SM staging half of the database tile: a third less L2->SM
// operand traffic, 6-deep ring).  Interleaved A/Bs on B200 (profiles/r01_session2_notes.md): with the lock-step
// window holding the wave together the pair form is 2-3.5 % faster for D = 768 one-term searches (C3: 1421-1442 vs
// 1385-1393 TFLOP/s) and ~10 % faster for the three-term split-precision search (C2); the query-stationary form
// (D <= 256) already stages database slices only and is faster as a single CTA.  Option "tc_cta_group" overrides.
int tc_cta_group(const rdb_handle* h, int nq, int nterms, int d) {
  if (nq <= TC_BM) return 1;
  if (h->opt.tc_cta_group == 1 || h->opt.tc_cta_group == 2) return h->opt.tc_cta_group;
  if (nterms == 3) return 2;
  return d > TcCfg<1>::ASTAT_MAX_KS * TC_BK ? 2 : 1;
}

// nqg = query-tile GROUPS (128 * cg queries each); S chunks of tiles_per_chunk 256-row database tiles
int launch_tc(rdb_handle* h, const void* qhi, const void* qlo, int nq, int k, int cg, int nqg, int S, int tiles_per_chunk,
              int ntiles, int nterms, float* ck, int* ci, int tile_step = 1, bool keep_gthr = false,
              const int* run_if = nullptr, float* dump = nullptr, long long dump_pitch = 0, int row_base = 0,
              int row_end = 0, const DevPlan* plan = nullptr, int share2 = 0) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = encode_2d(h, &p.tmap_q[0], qhi, nq, h->d, h->dp, TC_BM))) return rc;
  if (nterms == 3) { if ((rc = encode_2d(h, &p.tmap_q[1], qlo, nq, h->d, h->dp, TC_BM))) return rc; }
  else p.tmap_q[1] = p.tmap_q[0];
  p.ynorm = h->ynorm; p.ynmin32 = h->ynmin32; p.cand_key = ck; p.cand_idx = ci;
  const size_t gthr_bytes = size_t(nq) * 4 * (share2 ? 2 : 1);
  CUDA_TRY(h, h->gthr.ensure(gthr_bytes));
  if (!keep_gthr) CUDA_TRY(h, cudaMemsetAsync(h->gthr.p, 0, gthr_bytes, h->stream));
  p.share2 = (share2 && k <= 16) ? 1 : 0;
  p.tile_step = tile_step; p.run_if = run_if; p.plan = plan;
  p.nstages = std::max(2, h->opt.tc_stages);
  // norm slice (L2 keys straight from the accumulator): the queries of this search were staged as 2 q (search_impl)
  p.ext = (h->cur_hscale == 2.0f && h->yext) ? 1 : 0;
  if (p.ext) {
    if (h->qext_rows < nq) {
      const int64_t rows = round_up(std::max<int64_t>(nq, 4096), 4096);
      CUDA_TRY(h, h->qext.ensure(size_t(rows) * 16));
      qext_fill_kernel<<<unsigned((rows + 255) / 256), 256, 0, h->stream>>>(rows, h->qext.as<uint4>());
      h->launches++;
      h->qext_rows = rows;
    }
    if ((rc = encode_2d(h, &p.tmap_qx, h->qext.p, nq, 8, 8, TC_BM))) return rc;
  }
  p.astat = (nterms == 1 && (h->d + TC_BK - 1) / TC_BK + p.ext <= TcCfg<1>::ASTAT_MAX_KS && h->opt.tc_query_stationary) ? 1 : 0;
  p.gthr = (S > 1 || TC_LISTS > 1) ? h->gthr.as<uint32_t>() : nullptr;
  p.nq = nq; p.N = int(h->n); p.D = h->d;
  if (dump) { p.dump = dump; p.dump_pitch = dump_pitch; p.row_base = row_base; p.N = row_end; }   // k > 128: rows [row_base, row_end)
  p.nqt = nqg; p.S = S; p.tiles_per_chunk = tiles_per_chunk; p.ntiles = ntiles; p.kout = k;
  p.num_units = nqg * S; p.nterms = nterms;
#ifdef RDB_PROFILING
  p.dbg = h->opt.tc_debug;      // profiling builds only: skips the selection work, results invalid
#endif
  p.hint_q = kEvictLast;        // queries: re-read for every DB tile -> keep
  p.hint_y = kEvictNormal;      // database tiles: shared by the CTAs of a wave
  // lock-step window of the TMA producers (score_tc.cuh): on when there is more than one wave of units, at least a
  // quarter of a wave shares each chunk, and units are long enough to drift; option "tc_lockstep" = window in groups
  // of 8 tiles (0 = off).
  {
    const int ngroups = std::min(p.num_units, h->num_sms / cg);
    const int window = h->opt.tc_lockstep;
    const int sync_groups = (tiles_per_chunk + TC_SYNC_GS - 1) / TC_SYNC_GS;
    const int span = (ngroups + nqg - 1) / nqg + 1;                        // chunks one slot of ngroups units can touch
    // (measured: +6 % at C3, +4 % at the C5 shard, -4 % for the three-term split search -> one-term searches only)
    if (window > 0 && tile_step == 1 && nterms == 1 && p.num_units > ngroups && nqg * 2 >= ngroups &&
        sync_groups > 2 * window) {
      const size_t slots = size_t((p.num_units + ngroups - 1) / ngroups);
      const size_t bytes = (slots * span * size_t(sync_groups) + 1) * 4;  // + the "broken" flag
      CUDA_TRY(h, h->tcsync.ensure(bytes));
      CUDA_TRY(h, cudaMemsetAsync(h->tcsync.p, 0, bytes, h->stream));
      p.sync = h->tcsync.as<uint32_t>(); p.sync_groups = sync_groups; p.sync_window = window;
      p.sync_span = span;
      p.sync_broken = p.sync + slots * span * size_t(sync_groups);
      p.sync_spins = h->opt.tc_lockstep_spins;                            // default 4096: ~5 ms of patience
    }
  }
  rc = launch_tc_cg(h, p, k, cg);
  if (rc) return rc;
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

constexpr int64_t kQueryBatch = 65536;
constexpr int kMaxK = 128;          // fused selectors (register list / reservoir / streaming lists)
constexpr int kMaxKLarge = SELK_MAXK;   // 128 < k <= 2048: dense keys + radix select (run_largek); faiss-gpu's own limit
constexpr int64_t kQueryBatchLargeK = 4096;
// host-buffer searches at least this large upload their queries in pieces behind the running search (search_impl)
constexpr int64_t kPipeMinQueries = 4096;
constexpr size_t kPipeMinBytes = size_t(8) << 20;
constexpr int kLargeKQueryBlock = 256;                // queries per dense key block (two SIMT query tiles)
constexpr int64_t kLargeKRowsDefault = 1 << 20;       // rows per chunk: 256 x 1M x 4 B = 1 GiB of keys
constexpr int kMaxKTc = 128;       // k <= 32: register-resident list; 32 < k <= 128: local-memory reservoir
constexpr int kMaxKSplit = 104;    // split-precision path keeps kc = 16 / 32 / 64 / 128 candidates: slack >= 6 / 8 / 16 / 24
constexpr int64_t kMinRowsTc = 1024;
constexpr int kTier1Hold = 8;         // batches a raised tier-1 level is kept before it decays by one
constexpr int kTcSample = 64;          // large-k pivot: every 64th DB tile
constexpr int kTcPivotRank = 16;       // ... and the sample's 16th best key
constexpr int kTcPivotMinTiles = 1024; // >= 16 sampled tiles (N >= 262144)

// ---- small-batch streaming search (nq <= 4): ONE launch per pass does query prep + stream + final merge

// FILTER (sampled pivot) needs a sample that is a known fraction of the rows: the pivot pass takes every
// `sample_mul`-th warp step (each warp must own at least that many steps) and its rank-r key becomes the pivot.
// Expected rows above the pivot = r * sample_mul, chosen >= 4k (k of them exist with overwhelming probability) and
// << STREAM_FCAP; r <= 128 (what the LIST policies hold), so a sample of at most 1/4 of the rows is needed.
// warp steps per dynamically claimed row chunk of the streaming scorer: 4 (~48 KB per claim) once every warp gets at
// least four chunks, single steps on small shards so that the rows still spread over all warps of the grid
int stream_chunk_steps(rdb_handle* h, int blocks) {
  const int nvec = (h->store == RDB_STORE_F32) ? h->d / 4 : h->dp / 8;
  const int lpr_log2 = nvec >= 96 ? 5 : (nvec >= 48 ? 4 : 3);
  const int rw = ((h->store == RDB_STORE_F32) ? 4 : 8) * (32 >> lpr_log2);
  const int64_t steps = (h->n + rw - 1) / rw;
  return int(std::min<int64_t>(4, std::max<int64_t>(1, steps / (int64_t(blocks) * STREAM_WARPS * 4))));
}
int stream_rows_per_block(rdb_handle* h) {
  const int blocks0 = int(std::min<int64_t>(h->num_sms, (h->n + 31) / 32));
  return int(round_up((h->n + blocks0 - 1) / blocks0, 32));
}
int stream_sample_mul(rdb_handle* h, int rpb) {
  const int nvec = (h->store == RDB_STORE_F32) ? h->d / 4 : h->dp / 8;
  const int lpr_log2 = nvec >= 96 ? 5 : (nvec >= 48 ? 4 : 3);
  const int rw = ((h->store == RDB_STORE_F32) ? 4 : 8) * (32 >> lpr_log2);        // rows per warp step
  return std::min(STREAM_SAMPLE, rpb / (STREAM_WARPS * rw));
}
// rank of the sample whose key becomes the pivot: ~4k rows of the database should beat it (rank * sample_mul >= 4k)
int stream_pivot_rank(int k, int sample_mul) {
  return std::max(STREAM_PIVOT_RANK, (4 * k + sample_mul - 1) / std::max(sample_mul, 1));
}
bool stream_filter_ok(int k, int sample_mul) {
  return k > 32 && sample_mul >= 4 && stream_pivot_rank(k, sample_mul) <= 128;
}

// The whole nq <= 4 search.  Host buffers (the `predict()` latency path) never touch a copy engine when the queries fit
// the kernel parameters (nq * D <= STREAM_QINLINE floats): the queries ride in the launch itself, the kernel writes the
// packed results straight into mapped pinned host memory and publishes a sequence number there; the host spins on it.
// One launch, no cudaMemcpy, no stream synchronisation: ~10 us of host overhead instead of ~29 us.  Larger inputs use
// one pinned staging area: one H2D copy of the queries, the launches, ONE D2H copy of the packed results.
int search_stream(rdb_handle* h, const float* q, int nq, int k, int mem, int normalize, bool shard_mode, float* out_a,
                  int64_t* out_idx, float* out_lbl, float* out_qnorm) {
  const int D = h->d;
  const bool host = mem == RDB_MEM_HOST;
  cudaStream_t s = h->stream;
  const float* labels = (h->labels && h->nlabels == h->n) ? h->labels : nullptr;
  const int rpb = stream_rows_per_block(h);
  const int S = int((h->n + rpb - 1) / rpb);
  const int sample_mul = stream_sample_mul(h, rpb);
  const bool filter_ok = stream_filter_ok(k, sample_mul);
  const int mode = k <= 32 ? STREAM_LIST1 : (filter_ok ? STREAM_FILTER : STREAM_LIST4);
  const int pivot_rank = stream_pivot_rank(k, stream_sample_mul(h, stream_rows_per_block(h)));
  const int kc = std::max(k, pivot_rank);
  CUDA_TRY(h, h->cand_key.ensure(size_t(nq) * S * kc * 4));
  CUDA_TRY(h, h->cand_idx.ensure(size_t(nq) * S * kc * 4));
  if (!h->stream_ctl.p) {
    CUDA_TRY(h, h->stream_ctl.ensure(sizeof(StreamCtl)));
    CUDA_TRY(h, cudaMemsetAsync(h->stream_ctl.p, 0, sizeof(StreamCtl), s));
  }
  // packed outputs (host path): [a nq*k f32][lbl nq*k f32][qnorm 4 f32][idx nq*k i64][flag u32 ...]
  const size_t nk = size_t(nq) * k;
  const size_t off_l = nk * 4, off_qn = 2 * nk * 4, off_i = 2 * nk * 4 + 16, off_flag = off_i + nk * 8;
  const size_t pack_bytes = off_flag + 16;
  const size_t q_bytes = size_t(nq) * D * 4;
  const bool zero_copy = host && size_t(nq) * D <= size_t(STREAM_QINLINE);
  float* d_a = out_a; int64_t* d_i = out_idx; float* d_l = out_lbl; float* d_qn = out_qnorm;
  const float* qsrc = q;
  if (host) {
    if (h->pin_bytes < pack_bytes + q_bytes) {
      if (h->pin) cudaFreeHost(h->pin);
      h->pin = nullptr; h->pin_bytes = 0;
      const size_t want = std::max<size_t>(pack_bytes + q_bytes, 1 << 16);
      CUDA_TRY(h, cudaHostAlloc(&h->pin, want, cudaHostAllocMapped));
      h->pin_bytes = want;
      memset(h->pin, 0, want);
    }
    char* base;
    if (zero_copy) {
      // the kernel writes the packed results into the pinned block itself (unified addressing: same pointer)
      base = static_cast<char*>(h->pin);
      qsrc = nullptr;
    } else {
      CUDA_TRY(h, h->q_stage.ensure(q_bytes));
      CUDA_TRY(h, h->o_dist.ensure(pack_bytes));
      memcpy(static_cast<char*>(h->pin) + pack_bytes, q, q_bytes);
      CUDA_TRY(h, cudaMemcpyAsync(h->q_stage.p, static_cast<char*>(h->pin) + pack_bytes, q_bytes, cudaMemcpyHostToDevice, s));
      qsrc = h->q_stage.as<float>();
      base = static_cast<char*>(h->o_dist.p);
    }
    d_a = reinterpret_cast<float*>(base);
    d_l = out_lbl ? reinterpret_cast<float*>(base + off_l) : nullptr;
    d_qn = reinterpret_cast<float*>(base + off_qn);
    d_i = reinterpret_cast<int64_t*>(base + off_i);
  }
  // one StreamParams is reused for every launch of the search (it carries up to 8 KB of inline queries)
  StreamParams& p = h->stream_params;
  p.ynorm = h->ynorm; p.N = int(h->n);
  p.q_raw = qsrc; p.nq = nq; p.D = D; p.normalize = normalize; p.np = np_plan(h);
  if (zero_copy) memcpy(p.qin, q, q_bytes);
  p.rows_per_block = rpb; p.kout = k; p.step_mul = 1; p.chunk_steps = stream_chunk_steps(h, S);
  p.cand_key = h->cand_key.as<float>(); p.cand_idx = h->cand_idx.as<int>();
  p.fkey = nullptr; p.fidx = nullptr;
  p.ctl = h->stream_ctl.as<StreamCtl>();
  p.use_pivot_out = 0; p.run_if_fallback = 0;
  p.id_offset = h->id_offset; p.labels = labels;
  p.out_dist = shard_mode ? nullptr : d_a; p.out_key = shard_mode ? d_a : nullptr;
  p.out_idx = reinterpret_cast<long long*>(d_i); p.out_lbl = d_l; p.out_qnorm = d_qn;
  p.host_flag = nullptr; p.flag_seq = 0;
  p.prof = nullptr;
#ifdef RDB_PROFILING
  p.prof = reinterpret_cast<unsigned long long*>(h->opt.stream_prof);
#endif
  volatile unsigned int* flag = nullptr;
  unsigned int seq = 0;
  if (zero_copy) {
    flag = reinterpret_cast<volatile unsigned int*>(static_cast<char*>(h->pin) + off_flag);
    seq = ++h->stream_seq;
    if (seq == 0) seq = ++h->stream_seq;
    *flag = 0u;
  }
  int rc;
  cudaEventRecord(h->ev0, s);
  if (mode == STREAM_FILTER) {
    CUDA_TRY(h, h->fkey.ensure(size_t(4) * STREAM_FCAP * 4));
    CUDA_TRY(h, h->fidx.ensure(size_t(4) * STREAM_FCAP * 4));
    p.fkey = h->fkey.as<float>(); p.fidx = h->fidx.as<int>();
    // 1) pivot from a strided 1/64 sample (LIST, k = 16)
    float* sv_dist = p.out_dist; float* sv_key = p.out_key; long long* sv_idx = p.out_idx; float* sv_lbl = p.out_lbl;
    float* sv_qn = p.out_qnorm;
    p.kout = pivot_rank; p.step_mul = sample_mul; p.use_pivot_out = 1; p.chunk_steps = 1;
    p.out_dist = nullptr; p.out_key = nullptr; p.out_idx = nullptr; p.out_lbl = nullptr; p.out_qnorm = nullptr;
    if ((rc = launch_stream(h, p, S, pivot_rank <= 32 ? STREAM_LIST1 : STREAM_LIST4))) return rc;
    p.kout = k; p.step_mul = 1; p.use_pivot_out = 0; p.chunk_steps = stream_chunk_steps(h, S);
    p.out_dist = sv_dist; p.out_key = sv_key; p.out_idx = sv_idx; p.out_lbl = sv_lbl; p.out_qnorm = sv_qn;
    if ((rc = launch_stream(h, p, S, STREAM_FILTER))) return rc;      // 2) full pass: append key >= pivot, rank, emit
    p.run_if_fallback = 1;                     // 3) exits at once unless the FILTER pass raised the fallback flag
    p.host_flag = const_cast<unsigned int*>(flag); p.flag_seq = seq;   //    ... and publishes the results either way
    if ((rc = launch_stream(h, p, S, STREAM_LIST4))) return rc;
  } else {
    p.host_flag = const_cast<unsigned int*>(flag); p.flag_seq = seq;
    if ((rc = launch_stream(h, p, S, mode))) return rc;
  }
  cudaEventRecord(h->ev1, s);
  h->ev_valid = true; h->last_algo = RDB_ALGO_STREAM; h->last_S = S;
  if (host) {
    if (zero_copy) {
      // spin on the sequence number the last block publishes (bounded: fall back to the stream if nothing arrives)
      unsigned long long spins = 0;
      while (*flag != seq) {
        if ((++spins & 0xFFFFF) == 0) {                         // every ~1M polls: is the stream dead?
          const cudaError_t e = cudaStreamQuery(s);
          if (e == cudaSuccess) { if (*flag == seq) break; return fail(h, RDB_ERR_CUDA, "stream search: results never published"); }
          if (e != cudaErrorNotReady) { cudaGetLastError(); return fail(h, RDB_ERR_CUDA, std::string("stream search: ") + cudaGetErrorString(e)); }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      __atomic_thread_fence(__ATOMIC_ACQUIRE);
    } else {
      CUDA_TRY(h, cudaMemcpyAsync(h->pin, h->o_dist.p, pack_bytes, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, host_sync(h));
    }
    const char* base = static_cast<const char*>(h->pin);
    memcpy(out_a, base, nk * 4);
    if (out_lbl) memcpy(out_lbl, base + off_l, nk * 4);
    if (out_qnorm) memcpy(out_qnorm, base + off_qn, size_t(nq) * 4);
    memcpy(out_idx, base + off_i, nk * 8);
  }
  return RDB_OK;
}

// score + select over the local shard into h->cand_key / h->cand_idx; *L_out = lists per query (width kc each)
int run_scorer(rdb_handle* h, int algo, int nterms, const QueryView& qv, int kc, int* L_out, bool timed, int share2 = 0) {
  int rc, S, tpc;
  const int nqt = (qv.nq + 127) / 128;
  cudaStream_t s = h->stream;
  if (algo == RDB_ALGO_TC) {
    const int ntiles = int((h->n + TC_BN - 1) / TC_BN);
    const int cg = tc_cta_group(h, qv.nq, nterms, h->d);
    const int nqg = (qv.nq + TC_BM * cg - 1) / (TC_BM * cg);
    // large k: per-unit selection overhead (reservoir warm-up, final sort) is worth ~64 tiles -> fewer, longer units.
    // The minimum unit length is expressed in tiles of the C3 shape (12 K-slices, one term): long rows / three terms
    // make every tile proportionally longer, so small databases at the reference's D = 5376 still fill the machine.
    const int slices_per_tile = ((h->d + TC_BK - 1) / TC_BK) * nterms;
    const int base_min = kc > 32 ? 64 : 4;
    const int min_tiles = std::max(1, std::min(base_min, base_min * 12 / slices_per_tile));
    S = choose_splits(nqg, ntiles, h->num_sms / cg, 256 / TC_LISTS, min_tiles, &tpc,
                      (kc > 32 ? 64.0 : 2.0) * min_tiles / base_min);
    if (h->opt.tc_chunks > 0) {                                   // A/B option
      tpc = (ntiles + h->opt.tc_chunks - 1) / h->opt.tc_chunks;
      S = (ntiles + tpc - 1) / tpc;
    }
    CUDA_TRY(h, h->cand_key.ensure(size_t(qv.nq) * S * TC_LISTS * kc * 4));
    CUDA_TRY(h, h->cand_idx.ensure(size_t(qv.nq) * S * TC_LISTS * kc * 4));
    if (timed) cudaEventRecord(h->ev0, s);
    const bool pivoted = kc > 32 && nterms == 1 && ntiles >= kTcPivotMinTiles && h->opt.tc_pivot;
    if (pivoted) {
      // Large k: seed every query's admission bound from a strided 1/64 sample of the DB tiles (rank-16 key of the
      // sample: ~1000 rows of the shard beat it), so the reservoirs admit ~1000 rows per query in total instead of
      // k*ln(n/k) per work unit.  Exact whenever >= k rows pass; rdb checks that on the device (pivot_check below).
      const int ntiles_s = (ntiles + kTcSample - 1) / kTcSample;
      int tpc_s;
      const int S_s = choose_splits(nqg, ntiles_s, h->num_sms / cg, 256 / TC_LISTS, 1, &tpc_s);
      CUDA_TRY(h, h->cand_key.ensure(size_t(qv.nq) * S_s * TC_LISTS * kTcPivotRank * 4));
      CUDA_TRY(h, h->cand_idx.ensure(size_t(qv.nq) * S_s * TC_LISTS * kTcPivotRank * 4));
      CUDA_TRY(h, h->rr_key.ensure(size_t(qv.nq) * kTcPivotRank * 4));
      if ((rc = launch_tc(h, qv.qhi, qv.qlo, qv.nq, kTcPivotRank, cg, nqg, S_s, tpc_s, ntiles_s, 1,
                          h->cand_key.as<float>(), h->cand_idx.as<int>(), kTcSample))) return rc;
      {
        const int warps = 4;
        dim3 grid((qv.nq + warps - 1) / warps), block(32 * warps);
        merge_lists_kernel<int><<<grid, block, 0, s>>>(h->cand_key.as<float>(), h->cand_idx.as<int>(), nullptr, qv.nq,
                                                       S_s * TC_LISTS, kTcPivotRank, kTcPivotRank, 0, nullptr, 0, nullptr,
                                                       nullptr, nullptr, nullptr, h->rr_key.as<float>());
        gthr_from_pivot_kernel<<<(qv.nq + 255) / 256, 256, 0, s>>>(h->rr_key.as<float>(), qv.nq, kTcPivotRank,
                                                                   kTcPivotRank - 1, h->gthr.as<uint32_t>());
        h->launches += 2;
        CUDA_TRY(h, cudaGetLastError());
      }
    }
    if ((rc = launch_tc(h, qv.qhi, qv.qlo, qv.nq, kc, cg, nqg, S, tpc, ntiles, nterms, h->cand_key.as<float>(),
                        h->cand_idx.as<int>(), 1, pivoted, nullptr, nullptr, 0, 0, 0, nullptr, share2))) return rc;
    if (timed) cudaEventRecord(h->ev1, s);
    h->tc_pivoted = pivoted; h->tc_cg = cg; h->tc_nqg = nqg; h->tc_S = S; h->tc_tpc = tpc; h->tc_ntiles = ntiles;
    *L_out = S * TC_LISTS;
  } else {
    const int ntiles = int((h->n + SIMT_BN - 1) / SIMT_BN);
    S = choose_splits(nqt, ntiles, 2 * h->num_sms, 256 / SIMT_LISTS, 2, &tpc);
    CUDA_TRY(h, h->cand_key.ensure(size_t(qv.nq) * S * SIMT_LISTS * kc * 4));
    CUDA_TRY(h, h->cand_idx.ensure(size_t(qv.nq) * S * SIMT_LISTS * kc * 4));
    if (timed) cudaEventRecord(h->ev0, s);
    if ((rc = launch_simt(h, qv.qf, qv.qhi, qv.nq, kc, nqt, S, tpc * SIMT_BN, h->cand_key.as<float>(),
                          h->cand_idx.as<int>()))) return rc;
    if (timed) cudaEventRecord(h->ev1, s);
    *L_out = S * SIMT_LISTS;
  }
  if (timed) { h->ev_valid = true; h->last_algo = algo; h->last_S = S; }
  return RDB_OK;
}

// ---- large k (128 < k <= 2048): exact fp32 keys of (query block x row chunk) written to HBM by the CUDA-core scorer's
// DUMP form, exact radix select per query and chunk (select_large.cuh) -> one sorted list per chunk in
// h->cand_key / h->cand_idx, laid out [nq][S][k] for merge_lists_kernel.  *L_out = S.
bool largek_use_tc(const rdb_handle* h, int nterms) {
  bool use_tc = h->store != RDB_STORE_F32 && h->n >= kMinRowsTc;
  if (h->opt.largek_scorer == 1) use_tc = false;
  else if (h->opt.largek_scorer == 2 && h->store != RDB_STORE_F32 && h->n >= TC_BN) use_tc = true;
  return nterms == 3 ? true : use_tc;
}

int run_largek(rdb_handle* h, const QueryView& qv, int k, int* L_out, int nterms = 1) {
  const int64_t N = h->n;
  // 16-bit stores: the keys come from the tensor cores (SelectDump epilogue of kernel 2); fp32 stores need exact fp32
  // keys -> CUDA-core scorer.  Option "largek_scorer" = 1 (CUDA cores) | 2 (tensor cores) overrides (tests).
  // nterms == 3 (largek_split_search): approximate split-precision tensor-core keys of an fp32 store.
  const bool use_tc = largek_use_tc(h, nterms);
  const int64_t align = use_tc ? TC_BN : SIMT_BN;
  int64_t rows = kLargeKRowsDefault;
  if (h->opt.largek_rows > 0) rows = h->opt.largek_rows;       // tests: force several chunks
  rows = std::max<int64_t>(rows, (N + 255) / 256);          // merge_lists_kernel folds at most 256 lists
  rows = round_up(rows, align);
  rows = std::min<int64_t>(rows, round_up(N, align));
  const int S = int((N + rows - 1) / rows);
  const int qb = std::min(qv.nq, kLargeKQueryBlock);
  CUDA_TRY(h, h->lk_scores.ensure(size_t(qb) * size_t(rows) * 4));
  CUDA_TRY(h, h->cand_key.ensure(size_t(qv.nq) * S * k * 4));
  CUDA_TRY(h, h->cand_idx.ensure(size_t(qv.nq) * S * k * 4));
  float* scores = h->lk_scores.as<float>();
  cudaStream_t s = h->stream;
  CUDA_TRY(h, cudaFuncSetAttribute(select_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)selk_smem_bytes()));
  const int use_sample = h->opt.largek_sample ? 1 : 0;      // sampled-pivot fast path of the select (A/B option)
  cudaEventRecord(h->ev0, s);
  int rc;
  for (int q0 = 0; q0 < qv.nq; q0 += kLargeKQueryBlock) {
    const int nqs = std::min(kLargeKQueryBlock, qv.nq - q0);
    const int nqt = (nqs + SIMT_BM - 1) / SIMT_BM;
    for (int c = 0; c < S; ++c) {
      const int64_t row0 = int64_t(c) * rows, row_end = std::min<int64_t>(N, row0 + rows);
      const int len = int(row_end - row0);
      if (use_tc) {
        const int cg = tc_cta_group(h, nqs, nterms, h->d);
        const int nqg = (nqs + TC_BM * cg - 1) / (TC_BM * cg);
        const int tiles = (len + TC_BN - 1) / TC_BN;
        int tpc;
        const int units = choose_splits(nqg, tiles, h->num_sms / cg, 256, 4, &tpc);
        const char* qhi = static_cast<const char*>(qv.qhi) + size_t(q0) * h->dp * 2;
        const char* qlo = nterms == 3 ? static_cast<const char*>(qv.qlo) + size_t(q0) * h->dp * 2 : nullptr;
        if ((rc = launch_tc(h, qhi, qlo, nqs, k, cg, nqg, units, tpc, tiles, nterms, nullptr, nullptr, 1, false, nullptr,
                            scores, rows, int(row0), int(row_end)))) return rc;
      } else {
        const int tiles = (len + SIMT_BN - 1) / SIMT_BN;
        int units = std::min(tiles, std::max(1, 8 * h->num_sms / nqt));
        const int tpu = (tiles + units - 1) / units;
        units = (tiles + tpu - 1) / tpu;
        rc = launch_simt_dump(h, qv, q0, nqs, nqt, units, tpu * SIMT_BN, int(row0), int(row_end), scores, rows);
        if (rc) return rc;
      }
      select_dense_kernel<<<nqs, SELK_THREADS, selk_smem_bytes(), s>>>(scores, rows, len, int(row0), k, S, c, q0,
                                                                       use_sample, h->cand_key.as<float>(),
                                                                       h->cand_idx.as<int>());
      h->launches++;
      CUDA_TRY(h, cudaGetLastError());
    }
  }
  cudaEventRecord(h->ev1, s);
  h->ev_valid = true; h->last_algo = use_tc ? RDB_ALGO_TC : RDB_ALGO_SIMT; h->last_S = S;
  *L_out = S;
  return RDB_OK;
}

// fold the local candidate lists: final form (dist or key, global id, label)
constexpr int kMergeTreeMaxQueries = 2048;   // above this the one-warp-per-query merge has the better throughput
int run_merge_local(rdb_handle* h, int nq, int L, int kc, int kout, const float* qnorm, bool shard_mode, float* d_a,
                    int64_t* d_i, float* d_l, long long id_offset, const float* labels, float* raw_key,
                    const int* run_if = nullptr, const int* q_dev = nullptr, const int* l_dev = nullptr, int* sat = nullptr) {
  if (q_dev) {
    // device-sized launch (DevPlan): nq is the grid capacity, the real query / list counts are read on the device
    dim3 grid(std::min((nq + 3) / 4, h->num_sms * 8)), block(128);
    merge_lists_kernel<int, MERGE_LPL><<<grid, block, 0, h->stream>>>(
        h->cand_key.as<float>(), h->cand_idx.as<int>(), nullptr, nq, L, kc, kout, h->metric == RDB_METRIC_L2 ? 1 : 0, qnorm,
        id_offset, labels, shard_mode ? nullptr : d_a, reinterpret_cast<long long*>(d_i), d_l, shard_mode ? d_a : raw_key,
        nullptr, 0, q_dev, l_dev);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDB_OK;
  }
  if (L == 1 && kc == kout && !run_if) {
    // one sorted list per query (k > 128 with a single row chunk): elementwise conversion instead of k merge rounds
    const long long total = (long long)nq * kout;
    finalize_sorted_list_kernel<<<unsigned((total + 255) / 256), 256, 0, h->stream>>>(
        h->cand_key.as<float>(), h->cand_idx.as<int>(), nq, kout, h->metric == RDB_METRIC_L2 ? 1 : 0, qnorm, id_offset,
        labels, shard_mode ? nullptr : d_a, reinterpret_cast<long long*>(d_i), d_l, shard_mode ? d_a : raw_key);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDB_OK;
  }
  if (kc <= 32 && kout <= 32 && L >= 8 && nq <= kMergeTreeMaxQueries) {
    // few queries, many short lists: one block per query, lists folded by merge32 trees (merge.cuh)
    merge_lists_tree_kernel<<<nq, MERGE_TREE_WARPS * 32, 0, h->stream>>>(
        h->cand_key.as<float>(), h->cand_idx.as<int>(), nq, L, kc, kout, h->metric == RDB_METRIC_L2 ? 1 : 0, qnorm,
        id_offset, labels, shard_mode ? nullptr : d_a, reinterpret_cast<long long*>(d_i), d_l, shard_mode ? d_a : raw_key,
        run_if, sat);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDB_OK;
  }
  const size_t stage = merge_stage_bytes<int>(L, kc, kout);
  const int warps = stage > 12 * 1024 ? 2 : 4;             // <= 48 KB of dynamic shared memory per block
  dim3 grid((nq + warps - 1) / warps), block(32 * warps);
  // lists per lane: the per-round work of the merge is proportional to it, so use the smallest that holds L lists
#define MERGE_LOCAL(LPL)                                                                                         \
  merge_lists_kernel<int, LPL><<<grid, block, stage * warps, h->stream>>>(                                       \
      h->cand_key.as<float>(), h->cand_idx.as<int>(), nullptr, nq, L, kc, kout, h->metric == RDB_METRIC_L2 ? 1 : 0, \
      qnorm, id_offset, labels, shard_mode ? nullptr : d_a, reinterpret_cast<long long*>(d_i), d_l,              \
      shard_mode ? d_a : raw_key, run_if, int(stage), nullptr, nullptr, sat)
  if (L <= 32) MERGE_LOCAL(1); else if (L <= 64) MERGE_LOCAL(2); else if (L <= 128) MERGE_LOCAL(4); else MERGE_LOCAL(8);
#undef MERGE_LOCAL
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

// ---------------------------------------------------------------------------------------------- fp32 stores
// Exact-fp32 neighbours at tensor-core speed.  A *certified pass* scores the queries approximately on the tensor cores
// keeping kc > k candidates per query, re-scores the candidates exactly in fp32 (kernel 6) and certifies query q iff
//     exact_key[k-1] > approx_key_of_the_worst_candidate + B,     B = eps * |q| * max|y| (x 2 for the L2 key)
// -- then no row outside the candidate set can belong to the exact top-k.  Two approximations are used, cheapest first:
//   tier 1   q_hi.y_hi (ONE MMA term, bf16 roundings of both operands: eps = 2^-8 + 2^-18 + accumulation), kc = 32
//            (k <= 16; register-list epilogue) or 128 (k <= 64; reservoir epilogue + sampled admission bound).
//            A third of the tensor work and half of the database bytes of tier 2; certifies whenever the exact k-th
//            key clears the kc-th approximate key by ~0.4 % of |q||y| (N >= 262144 rows).
//   tier 2   q_lo.y_hi + q_hi.y_lo + q_hi.y_hi (three terms, eps = 3.02 * 2^-18 + accumulation), kc = 16 .. 128.
// Queries tier 1 cannot certify are compacted and go through tier 2; what tier 2 cannot certify (adversarial ties) is
// searched by the exact CUDA-core kernel.  Every tier ends in exact fp32 keys, so the result does not depend on which
// tier certified a query.
// Candidate-list capacity (lists of kc entries) of a device-sized stage over at most `cap` queries: 16 lists per query,
// and never fewer than what one query-tile group needs to be split into 128 chunks.
int64_t planned_lists_cap(int cap) { return std::max<int64_t>(round_up(cap, 256) * 16, int64_t(256) * 256); }

// `plan` == null: host-sized pass over v.nq queries.  `plan` != null: device-sized pass -- v.nq is the CAPACITY (grids,
// buffers, tensor maps), the real number of queries and the chunking live in *plan (written by plan_tc_kernel from a
// device-side count); every launch is enqueued unconditionally and returns at once when the plan holds no query.
int certified_pass(rdb_handle* h, const QueryView& v, int k, int kc, int nterms, bool shard_mode, float* o_a,
                   int64_t* o_i, float* o_l, const float* labels, int* ucount, int* ulist, bool timed,
                   const DevPlan* plan = nullptr, int list_k = 0) {
  // list_k != 0 (host-sized tier 1 only): two-list cover -- the scorer keeps lists of list_k = kc / 2 entries whose
  // union covers the best kc keys (SelectSmall<16, 2>, common.cuh); the merge flags the queries it cannot vouch for
  const int nb = v.nq, D = h->d;
  const bool l2 = h->metric == RDB_METRIC_L2;
  cudaStream_t s = h->stream;
  int rc, L = 0;
  const int* q_dev = plan ? &plan->nq : nullptr;
  const int lk = (list_k && !plan) ? list_k : kc;       // entries per candidate list
  int* sat = nullptr;
  if (lk != kc) {
    CUDA_TRY(h, h->sat.ensure(size_t(nb) * 4));
    sat = h->sat.as<int>();
  }
  if (!plan) {
    if ((rc = run_scorer(h, RDB_ALGO_TC, nterms, v, lk, &L, timed, lk != kc ? 1 : 0))) return rc;
  } else {
    const int ntiles = int((h->n + TC_BN - 1) / TC_BN);
    const int cg = nb > TC_BM ? 2 : 1;                       // must match plan_tc_kernel's cg (split3_search)
    const int64_t lists = planned_lists_cap(nb);
    CUDA_TRY(h, h->cand_key.ensure(size_t(lists) * kc * 4));
    CUDA_TRY(h, h->cand_idx.ensure(size_t(lists) * kc * 4));
    const int nqg_cap = (nb + TC_BM * cg - 1) / (TC_BM * cg);
    if ((rc = launch_tc(h, v.qhi, v.qlo, nb, kc, cg, nqg_cap, 256 / TC_LISTS, ntiles, ntiles, nterms, h->cand_key.as<float>(),
                        h->cand_idx.as<int>(), 1, false, nullptr, nullptr, 0, 0, 0, plan))) return rc;
    L = 256;
  }
  CUDA_TRY(h, h->rr_key.ensure(size_t(nb) * kc * 4));
  CUDA_TRY(h, h->rr_idx.ensure(size_t(nb) * kc * 8));
  CUDA_TRY(h, h->rr_key2.ensure(size_t(nb) * kc * 4));
  CUDA_TRY(h, h->rr_idx2.ensure(size_t(nb) * kc * 8));
  // approximate top-kc per query (local ids, raw keys)
  if ((rc = run_merge_local(h, nb, L, lk, kc, v.qnorm, false, nullptr, h->rr_idx.as<int64_t>(), nullptr, 0, nullptr,
                            h->rr_key.as<float>(), nullptr, q_dev, plan ? &plan->L : nullptr, sat))) return rc;
  CUDA_TRY(h, cudaMemsetAsync(ucount, 0, 4, s));
  const int nks = (D + TC_BK - 1) / TC_BK;
  const float n_mma = float(nks * (TC_BK / 16) * nterms);
  const float accum = 2.0f * (n_mma + 16.f) * 1.1920928955078125e-07f /*2^-23*/;
  // one-term pass of the main batch: measured rounding residuals tighten the certificate (see rerank_exact_kernel)
  const float* qres = (nterms == 1 && v.qf == h->qf.as<float>()) ? h->qres.as<float>() : nullptr;
  const float eps = (nterms == 3 ? 3.02f * 3.814697265625e-06f /*2^-18*/
                                 : 0.00390625f /*2 * 2^-9*/ + 3.814697265625e-06f /*2^-18*/) + accum;
  dim3 rgrid(plan ? std::min(nb, h->num_sms * 8) : nb), rblock(RERANK_THREADS);     // re-rank: one block per query
  // the kernel also writes the final form of the first k ranks (distances / global ids / labels, or raw keys for a shard)
  float* f_dist = shard_mode ? nullptr : o_a;
  float* f_key = shard_mode ? o_a : nullptr;
  if (l2) rerank_exact_kernel<true><<<rgrid, rblock, 0, s>>>(
      h->rr_idx.as<long long>(), h->rr_key.as<float>(), nb, kc, k, v.qf, h->master, h->ynorm, D, eps, v.qnorm,
      h->d_ynorm_max, h->n, h->rr_key2.as<float>(), h->rr_idx2.as<long long>(), ulist, ucount, qres, h->d_ynorm_max + 1,
      accum, q_dev, nullptr, f_dist, reinterpret_cast<long long*>(o_i), o_l, f_key, h->id_offset, labels, sat);
  else rerank_exact_kernel<false><<<rgrid, rblock, 0, s>>>(
      h->rr_idx.as<long long>(), h->rr_key.as<float>(), nb, kc, k, v.qf, h->master, h->ynorm, D, eps, v.qnorm,
      h->d_ynorm_max, h->n, h->rr_key2.as<float>(), h->rr_idx2.as<long long>(), ulist, ucount, qres, h->d_ynorm_max + 1,
      accum, q_dev, nullptr, f_dist, reinterpret_cast<long long*>(o_i), o_l, f_key, h->id_offset, labels, sat);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

// Counts of the certified search, read back WITHOUT blocking the stream: the device counters are copied to pinned host
// memory behind the batch and folded into the statistics / the adaptive tier-1 level when the copy has landed (checked
// at the next search, or waited for by the getters and by host-path searches, which synchronise anyway).
void counts_resolve(rdb_handle* h, bool wait) {
  if (!h->counts_pending) return;
  if (wait) cudaEventSynchronize(h->ev_counts);
  else if (cudaEventQuery(h->ev_counts) != cudaSuccess) { cudaGetLastError(); return; }
  h->counts_pending = false;
  const int m1 = h->pin_counts[0], m2 = h->pin_counts[1];
  if (h->pending_tier1) {
    h->last_tier1_queries += h->pending_nb;
    h->last_tier1_uncertified += m1;
    h->last_tier1_kc = h->pending_kc1;
    if (h->pending_kc1 < 128 ? (4 * int64_t(m1) > h->pending_nb) : (2 * int64_t(m1) > h->pending_nb)) {
      h->t1_level = h->pending_kc1 < 128 ? 1 : 2;
      h->t1_hold = kTier1Hold;
    }
  }
  h->last_uncertified += m2;
}
int counts_post(rdb_handle* h, const int* ucount1, const int* ucount2, bool tier1, int nb, int kc1) {
  counts_resolve(h, true);             // one set of counters in flight at a time
  if (!h->pin_counts) {
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->pin_counts), 16, cudaHostAllocDefault));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_counts, cudaEventDisableTiming));
  }
  h->pin_counts[0] = 0; h->pin_counts[1] = 0;
  if (ucount1) CUDA_TRY(h, cudaMemcpyAsync(&h->pin_counts[0], ucount1, 4, cudaMemcpyDeviceToHost, h->stream));
  if (ucount2) CUDA_TRY(h, cudaMemcpyAsync(&h->pin_counts[1], ucount2, 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev_counts, h->stream));
  h->counts_pending = true; h->pending_tier1 = tier1; h->pending_nb = nb; h->pending_kc1 = kc1;
  return RDB_OK;
}

// Tier 2 (three-term pass + exact re-rank + certificate) followed by the exact CUDA-core search of what it cannot
// certify; results [.][k] into o_*.  `count_dev` == null: the batch is v (v.nq queries, host-sized).  Otherwise v is a
// compacted sub-batch of CAPACITY v.nq whose real size is *count_dev (device), and tier 2 is device-sized as well.
// The exact fallback is ALWAYS device-sized: how many queries tier 2 refuses is only known on the device, and nothing
// here waits for it -- the whole search is stream-ordered (no host round trip, graph-capturable).
// *ucount2_out = device counter of the queries that took the exact fallback.
int split3_search(rdb_handle* h, const QueryView& v, int k, bool shard_mode, float* o_a, int64_t* o_i, float* o_l,
                  const float* labels, bool timed, const int* count_dev, const int** ucount2_out) {
  const int cap = v.nq, D = h->d;
  cudaStream_t s = h->stream;
  int rc;
  const int kc = (k <= 10) ? 16 : (k <= 24 ? 32 : (k <= 48 ? 64 : 128));
  CUDA_TRY(h, h->uncert.ensure(size_t(cap + 1) * 4));
  CUDA_TRY(h, h->dev_ctl.ensure(2 * sizeof(DevPlan)));
  int* ucount = h->uncert.as<int>();
  int* ulist = ucount + 1;
  DevPlan* plan2 = h->dev_ctl.as<DevPlan>();           // tier 2 (device-sized form only)
  DevPlan* plan3 = plan2 + 1;                          // exact fallback
  if (count_dev) {
    const int ntiles = int((h->n + TC_BN - 1) / TC_BN);
    const int cg = cap > TC_BM ? 2 : 1;
    const int slices_per_tile = ((D + TC_BK - 1) / TC_BK) * 3;
    const int base_min = kc > 32 ? 64 : 4;
    const int min_tiles = std::max(1, std::min(base_min, base_min * 12 / slices_per_tile));
    plan_tc_kernel<<<1, 32, 0, s>>>(count_dev, cap, cg, ntiles, h->num_sms / cg, 256 / TC_LISTS, min_tiles,
                                    float((kc > 32 ? 64.0 : 2.0) * min_tiles / base_min), planned_lists_cap(cap), plan2);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  if ((rc = certified_pass(h, v, k, kc, 3, shard_mode, o_a, o_i, o_l, labels, ucount, ulist, timed,
                           count_dev ? plan2 : nullptr))) return rc;
  // ---- exact CUDA-core search of the uncertified queries, device-sized
  const int ntiles_s = int((h->n + SIMT_BN - 1) / SIMT_BN);
  const int64_t lists3 = std::max<int64_t>(round_up(cap, 128) * 8, int64_t(128) * 256);
  plan_simt_kernel<<<1, 32, 0, s>>>(ucount, cap, ntiles_s, 2 * h->num_sms, 256 / SIMT_LISTS, lists3, plan3);
  h->launches++;
  CUDA_TRY(h, h->fb_qf.ensure(size_t(cap) * D * 4));
  CUDA_TRY(h, h->fb_qnorm.ensure(size_t(cap) * 4));
  CUDA_TRY(h, h->fb_a.ensure(size_t(cap) * k * 4));
  CUDA_TRY(h, h->fb_i.ensure(size_t(cap) * k * 8));
  CUDA_TRY(h, h->fb_l.ensure(size_t(cap) * k * 4));
  CUDA_TRY(h, h->cand_key.ensure(size_t(lists3) * k * 4));
  CUDA_TRY(h, h->cand_idx.ensure(size_t(lists3) * k * 4));
  const int gblocks = std::min((cap + 7) / 8, h->num_sms * 4);
  gather_f32_rows_kernel<<<gblocks, 256, 0, s>>>(v.qf, ulist, cap, D, h->fb_qf.as<float>(), &plan3->nq);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if ((rc = launch_ingest(h, h->fb_qf.as<float>(), cap, 0, 0, nullptr, nullptr, nullptr, h->fb_qnorm.as<float>(), 1.0f,
                          nullptr, &plan3->nq))) return rc;
  if ((rc = launch_simt(h, h->fb_qf.as<float>(), nullptr, cap, k, (cap + 127) / 128, 1, SIMT_BN, h->cand_key.as<float>(),
                        h->cand_idx.as<int>(), plan3))) return rc;
  if ((rc = run_merge_local(h, cap, 256, k, k, h->fb_qnorm.as<float>(), shard_mode, h->fb_a.as<float>(),
                            h->fb_i.as<int64_t>(), h->fb_l.as<float>(), h->id_offset, labels, nullptr, nullptr, &plan3->nq,
                            &plan3->s_L))) return rc;
  scatter_results_kernel<<<std::min((cap * k + 255) / 256, h->num_sms * 8), 256, 0, s>>>(
      ulist, cap, k, h->fb_a.as<float>(), h->fb_i.as<long long>(), h->fb_l.as<float>(), o_a,
      reinterpret_cast<long long*>(o_i), o_l, &plan3->nq);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if (ucount2_out) *ucount2_out = ucount;
  return RDB_OK;
}

constexpr int kTier1MaxK = 64;        // kc = 128 candidates: >= 2x slack (k = 64 vs 128: the gap is ~2.6 sd above the bound on Gaussian data)
constexpr int kTier1SmallK = 16;      // k <= 16 (the reference asks for top_k + 10 = 15): 32 candidates (register-list epilogue, no
                                      // sample pass); a failed query only costs its share of a three-term pass

int exact_split_search(rdb_handle* h, const QueryView& qv, int k, bool shard_mode, float* d_a, int64_t* d_i, float* d_l,
                       const float* labels) {
  const int nb = qv.nq, D = h->d, Dp = h->dp;
  cudaStream_t s = h->stream;
  int rc;
  const int64_t ntiles = (h->n + TC_BN - 1) / TC_BN;
  counts_resolve(h, false);            // fold the previous batch's counters in if they have arrived
  // Tier-1 level, adapted to how the data certifies: 0 = 32 candidates when k <= 16 (on iid Gaussian data at D = 768
  // the exact 10th key clears the 32nd approximate key by a wide margin under the measured-residual bound), 1 = 128
  // candidates (more than a quarter failed with 32: beyond that the three-term pass over the failures costs more than
  // the larger epilogue), 2 = no tier 1 (more than half failed with 128).  A raised level decays by one after
  // kTier1Hold batches, so a change of the data is picked up again.  The counters arrive asynchronously, so a level
  // change takes effect one batch late.
  if (h->t1_hold > 0 && --h->t1_hold == 0 && h->t1_level > 0) { h->t1_level--; h->t1_hold = h->t1_level > 0 ? kTier1Hold : 0; }
  // (Shards below 262 144 rows stay on the three-term pass: the 32-candidate form would run there, but with the few
  // query tiles such shards see, re-searching even a handful of uncertified queries costs another pass over the whole
  // database -- measured at the reference's own shapes, 25 423 x 5376 with 256 queries: 0.74 ms with tier 1 on clustered
  // data (19 of 256 queries uncertified), 0.48 ms without; profiles/r02_refscale_probe.jsonl.)
  const bool tier1 = k <= kTier1MaxK && ntiles >= kTcPivotMinTiles && h->t1_level < 2 && h->opt.tier1;
  const int* ucount2 = nullptr;
  if (!tier1) {
    if ((rc = split3_search(h, qv, k, shard_mode, d_a, d_i, d_l, labels, true, nullptr, &ucount2))) return rc;
    return counts_post(h, nullptr, ucount2, false, nb, 0);
  }
  int kc1 = (h->t1_level == 0 && k <= kTier1SmallK) ? 32 : 128;
  if (const int f = h->opt.tier1_kc)                      // A/B option: force the candidate count (never below what k needs)
    if (f == 32 || f == 64 || f == 128) kc1 = (k <= kTier1SmallK) ? f : 128;

  CUDA_TRY(h, h->uncert1.ensure(size_t(nb + 1) * 4));
  int* ucount = h->uncert1.as<int>();
  int* ulist = ucount + 1;
  // 32 candidates as a two-list cover of 16-entry lists: the epilogue keeps the register list of the k <= 16 searches
  // (C2: 12.65 -> ~11.4 ms scorer) -- option "tier1_share2" = 0 keeps 32-entry lists
  const int list_k = (kc1 == 32 && h->opt.tier1_share2) ? 16 : 0;
  if ((rc = certified_pass(h, qv, k, kc1, 1, shard_mode, d_a, d_i, d_l, labels, ucount, ulist, true, nullptr, list_k)))
    return rc;
  // ---- tier 2 on the queries tier 1 could not certify: compacted on the device, every launch device-sized from ucount
  CUDA_TRY(h, h->t2_qf.ensure(size_t(nb) * D * 4));
  CUDA_TRY(h, h->t2_qhi.ensure(size_t(nb) * Dp * 2));
  CUDA_TRY(h, h->t2_qlo.ensure(size_t(nb) * Dp * 2));
  CUDA_TRY(h, h->t2_qnorm.ensure(size_t(nb) * 4));
  CUDA_TRY(h, h->t2_a.ensure(size_t(nb) * k * 4));
  CUDA_TRY(h, h->t2_i.ensure(size_t(nb) * k * 8));
  CUDA_TRY(h, h->t2_l.ensure(size_t(nb) * k * 4));
  gather_f32_rows_kernel<<<std::min((nb + 7) / 8, h->num_sms * 4), 256, 0, s>>>(qv.qf, ulist, nb, D, h->t2_qf.as<float>(),
                                                                                ucount);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if ((rc = launch_ingest(h, h->t2_qf.as<float>(), nb, 0, 0, nullptr, h->t2_qhi.p, h->t2_qlo.p, h->t2_qnorm.as<float>(),
                          h->cur_hscale, nullptr, ucount))) return rc;
  QueryView sub{h->t2_qf.as<float>(), h->t2_qhi.p, h->t2_qlo.p, h->t2_qnorm.as<float>(), nb};
  if ((rc = split3_search(h, sub, k, shard_mode, h->t2_a.as<float>(), h->t2_i.as<int64_t>(), h->t2_l.as<float>(), labels,
                          false, ucount, &ucount2))) return rc;
  scatter_results_kernel<<<std::min((nb * k + 255) / 256, h->num_sms * 8), 256, 0, s>>>(
      ulist, nb, k, h->t2_a.as<float>(), h->t2_i.as<long long>(), h->t2_l.as<float>(), d_a,
      reinterpret_cast<long long*>(d_i), d_l, ucount);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return counts_post(h, ucount, ucount2, true, nb, kc1);
}

// ---- fp32 stores, k beyond the certified fused selectors (104 < k <= 2048): the same certificate idea on the dense-key
// path.  Split-precision tensor-core keys (three MMA terms, |error| <= B) are dumped to HBM, the radix select keeps
// kc = k + slack candidates per query, rerank_exact_large_kernel re-scores them exactly in fp32, sorts, and certifies
// query q iff exact_key[k-1] > worst approximate candidate key + B (no row outside the candidate set can then belong to
// the exact top-k).  Uncertified queries (ties at the boundary thicker than the slack) are searched by the exact
// CUDA-core dense-key path.  Replaces the 37 TFLOP/s FFMA scorer for these searches.
constexpr int kLargeKSlackMin = 32;
int largek_slack(int k) { return std::max(kLargeKSlackMin, k / 8); }

int largek_split_search(rdb_handle* h, const QueryView& qv, int k, bool shard_mode, float* d_a, int64_t* d_i, float* d_l,
                        const float* labels) {
  const int nb = qv.nq, D = h->d;
  const bool l2 = h->metric == RDB_METRIC_L2;
  cudaStream_t s = h->stream;
  const int kc = int(std::min<int64_t>(k + largek_slack(k), std::max<int64_t>(h->n, 1)));
  int rc, L = 0;
  if ((rc = run_largek(h, qv, kc, &L, 3))) return rc;
  // approximate top-kc per query: the per-chunk lists folded into one (local ids, raw keys)
  const float* ckey = h->cand_key.as<float>();
  const void* cidx = h->cand_idx.p;
  bool idx64 = false;
  if (L > 1) {
    CUDA_TRY(h, h->rr_key.ensure(size_t(nb) * kc * 4));
    CUDA_TRY(h, h->rr_idx.ensure(size_t(nb) * kc * 8));
    if ((rc = run_merge_local(h, nb, L, kc, kc, qv.qnorm, false, nullptr, h->rr_idx.as<int64_t>(), nullptr, 0, nullptr,
                              h->rr_key.as<float>()))) return rc;
    ckey = h->rr_key.as<float>(); cidx = h->rr_idx.p; idx64 = true;
  }
  CUDA_TRY(h, h->uncert.ensure(size_t(nb + 1) * 4));
  int* ucount = h->uncert.as<int>();
  int* ulist = ucount + 1;
  CUDA_TRY(h, cudaMemsetAsync(ucount, 0, 4, s));
  const int nks = (D + TC_BK - 1) / TC_BK;
  const float accum = 2.0f * (float(nks * (TC_BK / 16) * 3) + 16.f) * 1.1920928955078125e-07f /*2^-23*/;
  const float eps = 3.02f * 3.814697265625e-06f /*2^-18*/ + accum;
  int n2 = 1;
  while (n2 < kc) n2 <<= 1;
  const size_t smem = size_t(n2) * 8;
#define RERANK_LARGE(L2V, IDXT)                                                                                         \
  rerank_exact_large_kernel<L2V, IDXT><<<nb, RERANK_LARGE_THREADS, smem, s>>>(                                          \
      reinterpret_cast<const IDXT*>(cidx), ckey, nb, kc, k, qv.qf, h->master, h->ynorm, D, eps, qv.qnorm,               \
      h->d_ynorm_max, h->n, h->id_offset, labels, shard_mode ? nullptr : d_a, reinterpret_cast<long long*>(d_i), d_l,   \
      shard_mode ? d_a : nullptr, ulist, ucount)
  if (l2) { if (idx64) RERANK_LARGE(true, long long); else RERANK_LARGE(true, int); }
  else { if (idx64) RERANK_LARGE(false, long long); else RERANK_LARGE(false, int); }
#undef RERANK_LARGE
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  int m = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&m, ucount, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, host_sync(h));
  h->last_uncertified += m;
  if (m > 0) {
    // exact CUDA-core dense keys + radix select for the queries the certificate refused
    CUDA_TRY(h, h->fb_qf.ensure(size_t(m) * D * 4));
    CUDA_TRY(h, h->fb_qnorm.ensure(size_t(m) * 4));
    CUDA_TRY(h, h->fb_a.ensure(size_t(m) * k * 4));
    CUDA_TRY(h, h->fb_i.ensure(size_t(m) * k * 8));
    CUDA_TRY(h, h->fb_l.ensure(size_t(m) * k * 4));
    gather_f32_rows_kernel<<<(m + 7) / 8, 256, 0, s>>>(qv.qf, ulist, m, D, h->fb_qf.as<float>());
    h->launches++;
    if ((rc = launch_ingest(h, h->fb_qf.as<float>(), m, 0, 0, nullptr, nullptr, nullptr, h->fb_qnorm.as<float>())))
      return rc;
    QueryView fv{h->fb_qf.as<float>(), nullptr, nullptr, h->fb_qnorm.as<float>(), m};
    int Lf = 0;
    const int saved = h->opt.largek_scorer;
    h->opt.largek_scorer = 1;
    rc = run_largek(h, fv, k, &Lf, 1);
    h->opt.largek_scorer = saved;
    if (rc) return rc;
    if ((rc = run_merge_local(h, m, Lf, k, k, fv.qnorm, shard_mode, h->fb_a.as<float>(), h->fb_i.as<int64_t>(),
                              h->fb_l.as<float>(), h->id_offset, labels, nullptr))) return rc;
    scatter_results_kernel<<<(m * k + 255) / 256, 256, 0, s>>>(ulist, m, k, h->fb_a.as<float>(),
                                                               h->fb_i.as<long long>(), h->fb_l.as<float>(), d_a,
                                                               reinterpret_cast<long long*>(d_i), d_l);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  return RDB_OK;
}

// One search over the local shard.  shard_mode: out_a receives merge keys instead of distances.
int search_impl(rdb_handle* h, const float* q, int64_t nq, int k, int mem, int normalize, int algo, bool shard_mode,
                float* out_a, int64_t* out_idx, float* out_lbl, float* out_qnorm) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (nq < 0 || k < 1 || (!q && nq > 0) || !out_a || !out_idx)
    return fail(h, RDB_ERR_INVALID, "search: bad arguments (nq >= 0, k >= 1, non-null buffers)");
  if (k > kMaxKLarge) return fail(h, RDB_ERR_UNSUPPORTED, "search: k > 2048 is not supported (the limit of faiss-gpu itself)");
  bool largek = k > kMaxK;           // dense keys + radix select
  if (largek && algo == RDB_ALGO_STREAM)
    return fail(h, RDB_ERR_UNSUPPORTED, "search: the streaming scorer holds k <= 128 (use RDB_ALGO_AUTO for larger k)");
  if (nq == 0) return RDB_OK;
  const int D = h->d, Dp = h->dp;
  const bool host = mem == RDB_MEM_HOST;
  const bool sixteen = h->store != RDB_STORE_F32;
  // scorer selection.  16-bit stores: tcgen05 (1 term).  fp32 stores: tiered certified search on tcgen05 (one bf16
  // term first, then split precision with 3 terms) + exact fp32 re-rank + certificate, exact CUDA-core kernel for
  // whatever cannot be certified (and for small cases).  k > 128: dense keys + radix select (run_largek).
  const bool tc_ok = sixteen ? (k <= kMaxKTc && h->n >= TC_BN) : (k <= kMaxKSplit && h->n >= TC_BN);
  // small batches are a pure HBM stream of the stored rows: dedicated streaming scorer (exact fp32 for fp32 stores)
  const bool stream_ok = nq <= 4 && k <= 128 && h->n >= 1 && (sixteen || D % 4 == 0);
  // fp32 stores beyond the certified fused selectors (k > 104): split-precision tensor-core dense keys + exact re-rank
  // + certificate (largek_split_search) unless the caller forces the exact CUDA-core scorer
  const bool lk_split = !sixteen && k > kMaxKSplit && h->n >= kMinRowsTc && h->opt.largek_split && h->d % 4 == 0 &&
                        (algo == RDB_ALGO_AUTO || algo == RDB_ALGO_TC) && k + largek_slack(k) <= SELK_CAP / 2 &&
                        !(nq <= 4 && k <= kMaxK);      // tiny batches with k <= 128 stay on the streaming scorer
  if (lk_split) { largek = true; algo = RDB_ALGO_TC; }
  else if (largek) algo = RDB_ALGO_SIMT;
  if (algo == RDB_ALGO_AUTO) {
    algo = (stream_ok && h->n >= 4096) ? RDB_ALGO_STREAM : ((tc_ok && h->n >= kMinRowsTc) ? RDB_ALGO_TC : RDB_ALGO_SIMT);
  }
  if (algo == RDB_ALGO_STREAM && !stream_ok)
    return fail(h, RDB_ERR_UNSUPPORTED, "search: streaming scorer needs nq <= 4, k <= 128 (and D % 4 == 0 for fp32 stores)");
  if (algo == RDB_ALGO_TC && !tc_ok && !lk_split)
    return fail(h, RDB_ERR_UNSUPPORTED,
                "search: tensor-core scorer needs ntotal >= 256 and k <= 128 (16-bit store) / k <= 104 (fp32 store)");
  const bool split = (algo == RDB_ALGO_TC) && !sixteen;
  const float* labels = (h->labels && h->nlabels == h->n) ? h->labels : nullptr;
  cudaStream_t s = h->stream;
  counts_resolve(h, false);
  if (!h->counts_pending) { h->last_uncertified = 0; h->last_tier1_queries = 0; h->last_tier1_uncertified = 0; }
  // L2 on bf16 operands: when the tensor cores consume the 16-bit query copies they are staged as 2 q (norm slice)
  const bool tc_consumer = algo == RDB_ALGO_TC || (largek && sixteen && largek_use_tc(h, 1));
  h->cur_hscale = (tc_consumer && h->use_ext() && h->yext) ? 2.0f : 1.0f;
  if (algo == RDB_ALGO_STREAM)
    return search_stream(h, q, int(nq), k, mem, normalize, shard_mode, out_a, out_idx, out_lbl, out_qnorm);

  // Work list.  Device buffers: one piece per query batch.  HOST buffers: the batch is cut into a few pieces whose
  // host-to-device copies run on a second stream into two alternating staging buffers, so the (pageable) upload of
  // piece i + 1 overlaps the search of piece i -- C3: 201 MB of queries per step, ~15 ms that used to sit in front of
  // the scorer -- and the results of all pieces go back with one set of copies at the end.
  const int64_t qbatch = largek ? kQueryBatchLargeK : kQueryBatch;
  // (16-bit stores only: the certified fp32 search reads its counters back once per piece, which serialises the pieces
  // on the host -- measured 6 % slower at C2 -- and has nothing to gain from a cut.)
  const bool piped = host && !largek && !split && !out_qnorm && h->opt.host_pipeline && nq >= kPipeMinQueries &&
                     size_t(std::min<int64_t>(nq, qbatch)) * D * 4 >= kPipeMinBytes;
  if (piped && !h->copy_stream) {
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_stage_free[i], cudaEventDisableTiming));
    }
  }
  for (int64_t b0 = 0; b0 < nq; b0 += qbatch) {
    const int nb = int(std::min<int64_t>(qbatch, nq - b0));
    // ---- pieces of this batch (offset, count)
    std::vector<std::pair<int, int>> pieces;
    if (piped && nb >= kPipeMinQueries) {
      // Copy-bound searches (small shards) take four equal pieces (the last piece's search is what stays exposed).
      // Compute-bound ones (large shards) want a small first piece -- its upload is the only one exposed -- and as few
      // pieces as possible after it, because every extra launch of the persistent scorer pays its own ramp and tail
      // (C3 cut 4096 / 16384 / 45056 lost as much as the overlap won): piece i + 1 may be r times piece i, r = how many
      // times faster a query uploads than it is searched (~ rows / 0.25 M at D = 768-class shapes, halved for safety),
      // two pieces when r >= 8, three otherwise.
      const bool compute_bound = h->n >= 400000;
      const double r = std::min(32.0, std::max(2.0, double(h->n) / 0.5e6));
      const double parts = r >= 8.0 ? 1.0 + r : 1.0 + r + r * r;
      int p0 = compute_bound ? std::max(1024, int(double(nb) / parts) / 256 * 256) : std::max(1024, (nb / 4 + 255) / 256 * 256);
      int off = 0, cur = p0;
      while (off < nb) {
        int take = std::min(cur, nb - off);
        // no small tail piece (compute-bound C3: 3072 + 62464, not ... + 1024: a 1024-query launch of the scorer is inefficient)
        if (nb - off - take < (compute_bound ? std::max(4096, take / 4) : 1024)) take = nb - off;
        pieces.emplace_back(off, take);
        off += take;
        if (compute_bound) cur = int(std::min<double>(double(cur) * r, double(nb)));
      }
    } else {
      pieces.emplace_back(0, nb);
    }
    const bool overlap = pieces.size() > 1;
    int max_piece = 0;
    for (auto& pc : pieces) max_piece = std::max(max_piece, pc.second);
    // ---- output views (device scratch for the whole batch when the caller's buffers are on the host)
    float* b_a = out_a + b0 * k;
    int64_t* b_i = out_idx + b0 * k;
    float* b_l = out_lbl ? out_lbl + b0 * k : nullptr;
    if (host) {
      CUDA_TRY(h, h->o_dist.ensure(size_t(nb) * k * 4));
      CUDA_TRY(h, h->o_idx.ensure(size_t(nb) * k * 8));
      if (out_lbl) CUDA_TRY(h, h->o_lbl.ensure(size_t(nb) * k * 4));
      b_a = h->o_dist.as<float>(); b_i = h->o_idx.as<int64_t>(); b_l = out_lbl ? h->o_lbl.as<float>() : nullptr;
      CUDA_TRY(h, h->q_stage.ensure(size_t(max_piece) * D * 4));
      if (overlap) CUDA_TRY(h, h->q_stage2.ensure(size_t(max_piece) * D * 4));
    }
    // upload of piece i (host path): plain copy on the search stream, or -- overlapped form -- on the copy stream into
    // staging buffer i & 1 once the query prep that last read that buffer is done
    auto stage_of = [&](size_t i) { return (overlap && (i & 1)) ? h->q_stage2.as<float>() : h->q_stage.as<float>(); };
    auto upload = [&](size_t i) -> cudaError_t {
      const float* src = q + (b0 + pieces[i].first) * D;
      const size_t bytes = size_t(pieces[i].second) * D * 4;
      if (!overlap) return cudaMemcpyAsync(stage_of(i), src, bytes, cudaMemcpyHostToDevice, s);
      cudaError_t e = cudaSuccess;
      if (i >= 2) e = cudaStreamWaitEvent(h->copy_stream, h->ev_stage_free[i & 1], 0);
      else if (i == 0) {
        // the staging buffers may still be read by work queued earlier on the search stream
        e = cudaEventRecord(h->ev_stage_free[0], s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->copy_stream, h->ev_stage_free[0], 0);
      }
      if (e == cudaSuccess) e = cudaMemcpyAsync(stage_of(i), src, bytes, cudaMemcpyHostToDevice, h->copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(h->ev_h2d[i & 1], h->copy_stream);
      return e;
    };
    if (host) CUDA_TRY(h, upload(0));
    for (size_t pi = 0; pi < pieces.size(); ++pi) {
    const int p_off = pieces[pi].first, pn = pieces[pi].second;
    const float* qsrc = q + (b0 + p_off) * D;
    if (host) {
      qsrc = stage_of(pi);
      if (overlap) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_h2d[pi & 1], 0));
    }
    // ---- query prep: normalise, |q|^2, convert (same fused kernel as ingest)
    CUDA_TRY(h, h->qnorm.ensure(size_t(pn) * 4));
    int rc;
    QueryView qv{nullptr, nullptr, nullptr, h->qnorm.as<float>(), pn};
    if (sixteen) {
      CUDA_TRY(h, h->qhi.ensure(size_t(pn) * Dp * 2));
      if ((rc = launch_ingest(h, qsrc, pn, normalize, 1, nullptr, h->qhi.p, nullptr, h->qnorm.as<float>(), h->cur_hscale)))
        return rc;
      qv.qhi = h->qhi.p;
    } else {
      CUDA_TRY(h, h->qf.ensure(size_t(pn) * D * 4));
      if (split) {
        CUDA_TRY(h, h->qhi.ensure(size_t(pn) * Dp * 2));
        CUDA_TRY(h, h->qlo.ensure(size_t(pn) * Dp * 2));
        CUDA_TRY(h, h->qres.ensure(size_t(pn) * 4));
      }
      if ((rc = launch_ingest(h, qsrc, pn, normalize, 0, h->qf.as<float>(), split ? h->qhi.p : nullptr,
                              split ? h->qlo.p : nullptr, h->qnorm.as<float>(), h->cur_hscale,
                              split ? h->qres.as<float>() : nullptr))) return rc;
      qv.qf = h->qf.as<float>(); qv.qhi = h->qhi.p; qv.qlo = h->qlo.p;
    }
    if (overlap) CUDA_TRY(h, cudaEventRecord(h->ev_stage_free[pi & 1], s));   // the prep was the staging buffer's last reader
    float* d_a = b_a + size_t(p_off) * k;
    int64_t* d_i = b_i + size_t(p_off) * k;
    float* d_l = b_l ? b_l + size_t(p_off) * k : nullptr;
    int L = 0;
    if (!split) {
      // ---- score + select, then merge
      h->tc_pivoted = false;
      if (h->n > 0 && (rc = largek ? run_largek(h, qv, k, &L) : run_scorer(h, algo, 1, qv, k, &L, true))) return rc;
      if ((rc = run_merge_local(h, pn, L, k, k, qv.qnorm, shard_mode, d_a, d_i, d_l, h->id_offset, labels, nullptr)))
        return rc;
      if (h->tc_pivoted) {
        // every query must have found min(k, n) rows above its sampled pivot; otherwise (flag) the two launches below
        // redo the batch without any bound -- they exit at once when the flag is clear (the normal case)
        CUDA_TRY(h, h->uncert.ensure(8));
        int* flag = h->uncert.as<int>();
        CUDA_TRY(h, cudaMemsetAsync(flag, 0, 4, s));
        check_complete_kernel<<<(pn + 255) / 256, 256, 0, s>>>(reinterpret_cast<const long long*>(d_i), pn, k,
                                                              int(std::min<int64_t>(k, h->n)), h->gthr.as<uint32_t>(), flag);
        h->launches++;
        if ((rc = launch_tc(h, qv.qhi, qv.qlo, pn, k, h->tc_cg, h->tc_nqg, h->tc_S, h->tc_tpc, h->tc_ntiles, 1,
                            h->cand_key.as<float>(), h->cand_idx.as<int>(), 1, true, flag))) return rc;
        if ((rc = run_merge_local(h, pn, L, k, k, qv.qnorm, shard_mode, d_a, d_i, d_l, h->id_offset, labels, nullptr,
                                  flag))) return rc;
      }
    } else {
      // ---- fp32 store on the tensor cores: certified approximate pass(es) + exact fp32 re-rank
      if ((rc = lk_split ? largek_split_search(h, qv, k, shard_mode, d_a, d_i, d_l, labels)
                         : exact_split_search(h, qv, k, shard_mode, d_a, d_i, d_l, labels))) return rc;
    }
    // the next piece's upload is issued BEHIND this piece's launches: the host blocks in it (pageable memory is staged
    // by the driver) while the GPU searches
    if (host && pi + 1 < pieces.size()) CUDA_TRY(h, upload(pi + 1));
    }  // pieces
    if (out_qnorm)
      CUDA_TRY(h, cudaMemcpyAsync(out_qnorm + b0, h->qnorm.p, size_t(nb) * 4,
                                  host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
    if (host) {
      CUDA_TRY(h, cudaMemcpyAsync(out_a + b0 * k, b_a, size_t(nb) * k * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, cudaMemcpyAsync(out_idx + b0 * k, b_i, size_t(nb) * k * 8, cudaMemcpyDeviceToHost, s));
      if (out_lbl) CUDA_TRY(h, cudaMemcpyAsync(out_lbl + b0 * k, b_l, size_t(nb) * k * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, host_sync(h));   // scratch is reused by the next batch
      counts_resolve(h, true);
    }
  }
  return RDB_OK;
}

// stream-ordered copy by the SMs of the launching device (rdb_copy_async): either side may be peer memory
template <typename V>
__global__ void __launch_bounds__(256) copy_bytes_kernel(V* __restrict__ dst, const V* __restrict__ src, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// every grow-only search scratch buffer of a handle (released by rdb_release_scratch / rdb_destroy, counted by rdb_mem_info)
using DevBufMember = DevBuf rdb_handle::*;
const DevBufMember kScratch[] = {
    &rdb_handle::add_stage, &rdb_handle::q_stage, &rdb_handle::q_stage2, &rdb_handle::qf, &rdb_handle::qhi, &rdb_handle::qlo, &rdb_handle::qnorm,
    &rdb_handle::cand_key, &rdb_handle::cand_idx, &rdb_handle::o_dist, &rdb_handle::o_idx, &rdb_handle::o_lbl,
    &rdb_handle::ids_stage, &rdb_handle::rec_stage, &rdb_handle::rr_key, &rdb_handle::rr_idx, &rdb_handle::rr_key2,
    &rdb_handle::rr_idx2, &rdb_handle::uncert, &rdb_handle::fb_qf, &rdb_handle::fb_qnorm, &rdb_handle::fb_a,
    &rdb_handle::fb_i, &rdb_handle::fb_l, &rdb_handle::gthr, &rdb_handle::tcsync, &rdb_handle::stream_ctl,
    &rdb_handle::fkey, &rdb_handle::fidx, &rdb_handle::lk_scores, &rdb_handle::uncert1, &rdb_handle::t2_qf,
    &rdb_handle::t2_qhi, &rdb_handle::t2_qlo, &rdb_handle::t2_qnorm, &rdb_handle::t2_a, &rdb_handle::t2_i,
    &rdb_handle::t2_l, &rdb_handle::dev_ctl, &rdb_handle::qext, &rdb_handle::qres, &rdb_handle::sat,
    &rdb_handle::res_stage};

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int rdb_abi_version(void) { return 2; }

const char* rdb_last_error(rdb_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int rdb_create(int d, int metric, int store_dtype, int device, unsigned flags, rdb_handle** out) {
  if (!out) return fail(nullptr, RDB_ERR_INVALID, "rdb_create: out is null");
  *out = nullptr;
  if (d < 1 || d > (1 << 20)) return fail(nullptr, RDB_ERR_INVALID, "rdb_create: dimension out of range");
  if (metric != RDB_METRIC_L2 && metric != RDB_METRIC_IP) return fail(nullptr, RDB_ERR_INVALID, "rdb_create: unknown metric");
  if (store_dtype < RDB_STORE_F32 || store_dtype > RDB_STORE_F16) return fail(nullptr, RDB_ERR_INVALID, "rdb_create: unknown store dtype");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, RDB_ERR_CUDA, "rdb_create: no CUDA device available (this index has no CPU fallback)");
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= ndev) return fail(nullptr, RDB_ERR_INVALID, "rdb_create: device ordinal out of range");
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, RDB_ERR_UNSUPPORTED, std::string("rdb_create: built for sm_100a (B200); device is sm_") +
                                                  std::to_string(prop.major) + std::to_string(prop.minor));
  rdb_handle* h = new rdb_handle();
  h->d = d; h->dp = int(round_up(d, 8)); h->metric = metric; h->store = store_dtype; h->device = device; h->flags = flags;
  h->num_sms = prop.multiProcessorCount;
  DeviceGuard dg(device);
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e != cudaSuccess) { delete h; cudaGetLastError(); return fail(nullptr, RDB_ERR_CUDA, std::string("rdb_create: ") + cudaGetErrorString(e)); }
  h->stream = h->own_stream;
  if (cudaMalloc(&h->d_ynorm_max, 8) != cudaSuccess || cudaMemset(h->d_ynorm_max, 0, 8) != cudaSuccess) {
    cudaGetLastError();
    rdb_destroy(h);
    return fail(nullptr, RDB_ERR_NOMEM, "rdb_create: device allocation failed");
  }
  if (np_plan_build(h) != RDB_OK) { rdb_destroy(h); return fail(nullptr, RDB_ERR_NOMEM, "rdb_create: device allocation failed"); }
  *out = h;
  return RDB_OK;
}

int rdb_destroy(rdb_handle* h) {
  if (!h) return RDB_OK;
  {
    DeviceGuard dg(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->master); cudaFree(h->hi); cudaFree(h->lo); cudaFree(h->ynorm); cudaFree(h->ynmin32); cudaFree(h->yext); cudaFree(h->labels);
    cudaFree(h->d_ynorm_max);
    for (auto m : kScratch) (h->*m).release();
    h->np_tab.release();
    if (h->pin) cudaFreeHost(h->pin);
    if (h->pin_counts) cudaFreeHost(h->pin_counts);
    if (h->ev_counts) cudaEventDestroy(h->ev_counts);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < 2; ++i) {
      if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
      if (h->ev_stage_free[i]) cudaEventDestroy(h->ev_stage_free[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    cudaGetLastError();
  }
  delete h;
  return RDB_OK;
}

int rdb_set_stream(rdb_handle* h, void* cuda_stream) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  cudaStreamSynchronize(h->stream);
  h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);   // NULL == legacy default stream
  return RDB_OK;
}

int rdb_use_own_stream(rdb_handle* h) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  cudaStreamSynchronize(h->stream);
  h->stream = h->own_stream;
  return RDB_OK;
}

int rdb_sync(rdb_handle* h) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  DeviceGuard dg(h->device);
  CUDA_TRY(h, host_sync(h));
  return RDB_OK;
}

int rdb_reserve(rdb_handle* h, int64_t n_total) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  return grow_to(h, n_total, /*exact=*/true);
}

int rdb_add(rdb_handle* h, const float* x, int64_t n, int mem, int normalize) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (n < 0 || (n > 0 && !x)) return fail(h, RDB_ERR_INVALID, "add: bad arguments");
  if (n == 0) return RDB_OK;
  int rc = grow_to(h, h->n + n);
  if (rc) return rc;
  const size_t D = h->d, Dp = h->dp;
  const int es = 2;
  const int norm_of_hi = h->store != RDB_STORE_F32;
  // host input is staged in slices so the staging buffer stays bounded (256 MiB)
  const int64_t slice = (mem == RDB_MEM_HOST) ? std::max<int64_t>(1, (int64_t(256) << 20) / int64_t(D * 4)) : n;
  for (int64_t s0 = 0; s0 < n; s0 += slice) {
    const int64_t m = std::min(slice, n - s0);
    const float* src = x + s0 * D;
    if (mem == RDB_MEM_HOST) {
      CUDA_TRY(h, h->add_stage.ensure(size_t(m) * D * 4));
      CUDA_TRY(h, cudaMemcpyAsync(h->add_stage.p, src, size_t(m) * D * 4, cudaMemcpyHostToDevice, h->stream));
      src = h->add_stage.as<float>();
    }
    const int64_t row0 = h->n + s0;
    float* master = h->has_master() ? h->master + row0 * D : nullptr;
    void* hi = reinterpret_cast<char*>(h->hi) + size_t(row0) * Dp * es;
    void* lo = h->has_lo() ? reinterpret_cast<char*>(h->lo) + size_t(row0) * Dp * es : nullptr;
    float* res2 = nullptr;
    if (lo) { CUDA_TRY(h, h->res_stage.ensure(size_t(m) * 4)); res2 = h->res_stage.as<float>(); }
    if ((rc = launch_ingest(h, src, m, normalize, norm_of_hi, master, hi, lo, h->ynorm + row0, 1.0f, res2))) return rc;
    ynorm_max_kernel<<<unsigned(std::min<int64_t>((m + 255) / 256, 1024)), 256, 0, h->stream>>>(h->ynorm + row0, m,
                                                                                            h->d_ynorm_max);
    if (res2) {
      ynorm_max_kernel<<<unsigned(std::min<int64_t>((m + 255) / 256, 1024)), 256, 0, h->stream>>>(res2, m, h->d_ynorm_max + 1);
      h->launches++;
    }
    {
      // minima of the (aligned) 32-row groups these rows touch; unfilled rows of the last group still hold |y|^2 = 0,
      // which only loosens the bound until they are added
      const int64_t g0 = row0 / 32, g1 = (row0 + m - 1) / 32;
      if (h->yext)     // bf16 operands: the L2 key comes out of the accumulator (norm slice), no epilogue filter needed
        yext_fill_kernel<<<unsigned((m + 255) / 256), 256, 0, h->stream>>>(h->ynorm + row0, m,
                                                                            static_cast<uint4*>(h->yext) + row0);
      else
        ynorm_min32_kernel<<<unsigned((g1 - g0 + 1 + 7) / 8), 256, 0, h->stream>>>(h->ynorm, g0, g1 - g0 + 1, h->ynmin32);
    }
    h->launches += 2;
    if (mem == RDB_MEM_HOST) CUDA_TRY(h, host_sync(h));  // staging buffer reuse
  }
  h->n += n;
  return RDB_OK;
}

int rdb_search(rdb_handle* h, const float* q, int64_t nq, int k, int mem, int normalize, float* out_dist,
               int64_t* out_idx, float* out_labels) {
  return search_impl(h, q, nq, k, mem, normalize, RDB_ALGO_AUTO, false, out_dist, out_idx, out_labels, nullptr);
}

int rdb_search_algo(rdb_handle* h, const float* q, int64_t nq, int k, int mem, int normalize, int algo,
                    float* out_dist, int64_t* out_idx, float* out_labels) {
  return search_impl(h, q, nq, k, mem, normalize, algo, false, out_dist, out_idx, out_labels, nullptr);
}

int rdb_search_shard(rdb_handle* h, const float* q_dev, int64_t nq, int k, int normalize, float* out_key,
                     int64_t* out_idx, float* out_labels, float* out_qnorm) {
  return search_impl(h, q_dev, nq, k, RDB_MEM_DEVICE, normalize, RDB_ALGO_AUTO, true, out_key, out_idx, out_labels,
                     out_qnorm);
}

int rdb_merge_shards(rdb_handle* h, const float* key, const int64_t* idx, const float* labels, int64_t nq, int nlists,
                     int k, const float* qnorm, float* out_dist, int64_t* out_idx, float* out_labels) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (nq < 0 || k < 1 || nlists < 1 || nlists > 32 * MERGE_LPL || !key || !idx || !out_dist || !out_idx)
    return fail(h, RDB_ERR_INVALID, "merge_shards: bad arguments");
  if (h->metric == RDB_METRIC_L2 && !qnorm) return fail(h, RDB_ERR_INVALID, "merge_shards: L2 needs qnorm");
  if (nq == 0) return RDB_OK;
  const int warps = 4;
  dim3 grid(unsigned((nq + warps - 1) / warps)), block(32 * warps);
  merge_lists_kernel<long long><<<grid, block, 0, h->stream>>>(
      key, reinterpret_cast<const long long*>(idx), labels, int(nq), nlists, k, k, h->metric == RDB_METRIC_L2 ? 1 : 0,
      qnorm, 0, nullptr, out_dist, reinterpret_cast<long long*>(out_idx), out_labels, nullptr);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

int rdb_ipc_alloc(rdb_handle* h, size_t bytes, void** dev_ptr, unsigned char* handle_out) {
  if (!h || !dev_ptr || !handle_out || bytes == 0) return fail(h, RDB_ERR_INVALID, "ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  void* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t mh;
  cudaError_t e = cudaIpcGetMemHandle(&mh, p);
  if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return fail(h, RDB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  memcpy(handle_out, &mh, 64);
  *dev_ptr = p;
  return RDB_OK;
}

int rdb_ipc_open(rdb_handle* h, const unsigned char* handle, void** dev_ptr) {
  if (!h || !handle || !dev_ptr) return fail(h, RDB_ERR_INVALID, "ipc_open: bad arguments");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle, 64);
  CUDA_TRY(h, cudaIpcOpenMemHandle(dev_ptr, mh, cudaIpcMemLazyEnablePeerAccess));
  return RDB_OK;
}

int rdb_ipc_close(rdb_handle* h, void* dev_ptr) {
  if (!h || !dev_ptr) return fail(h, RDB_ERR_INVALID, "ipc_close: bad arguments");
  DeviceGuard dg(h->device);
  CUDA_TRY(h, cudaIpcCloseMemHandle(dev_ptr));
  return RDB_OK;
}

int rdb_ipc_free(rdb_handle* h, void* dev_ptr) {
  if (!h || !dev_ptr) return fail(h, RDB_ERR_INVALID, "ipc_free: bad arguments");
  DeviceGuard dg(h->device);
  cudaStreamSynchronize(h->stream);
  CUDA_TRY(h, cudaFree(dev_ptr));
  return RDB_OK;
}

int rdb_merge_shards_peer(rdb_handle* h, const void* const* key_ptrs, const void* const* idx_ptrs,
                          const void* const* lbl_ptrs, int nlists, int64_t nq, int k, const float* qnorm,
                          float* out_dist, int64_t* out_idx, float* out_labels) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (nq < 0 || k < 1 || nlists < 1 || nlists > 32 || !key_ptrs || !idx_ptrs || !out_dist || !out_idx)
    return fail(h, RDB_ERR_INVALID, "merge_shards_peer: bad arguments (1 <= nlists <= 32)");
  if (h->metric == RDB_METRIC_L2 && !qnorm) return fail(h, RDB_ERR_INVALID, "merge_shards_peer: L2 needs qnorm");
  if (nq == 0) return RDB_OK;
  PeerLists P;
  memset(&P, 0, sizeof(P));
  for (int g = 0; g < nlists; ++g) {
    P.key[g] = reinterpret_cast<const float*>(key_ptrs[g]);
    P.idx[g] = reinterpret_cast<const long long*>(idx_ptrs[g]);
    P.lbl[g] = lbl_ptrs ? reinterpret_cast<const float*>(lbl_ptrs[g]) : nullptr;
  }
  const int warps = 4;
  dim3 grid(unsigned((nq + warps - 1) / warps)), block(32 * warps);
  merge_peer_lists_kernel<<<grid, block, 0, h->stream>>>(P, nlists, int(nq), k, h->metric == RDB_METRIC_L2 ? 1 : 0,
                                                         qnorm, out_dist, reinterpret_cast<long long*>(out_idx),
                                                         out_labels);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

int rdb_reconstruct_batch(rdb_handle* h, const int64_t* ids, int64_t n, int mem, float* out) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (n < 0 || (n > 0 && (!ids || !out))) return fail(h, RDB_ERR_INVALID, "reconstruct_batch: bad arguments");
  if (n == 0) return RDB_OK;
  const size_t D = h->d;
  const long long* d_ids = reinterpret_cast<const long long*>(ids);
  float* d_out = out;
  if (mem == RDB_MEM_HOST) {
    CUDA_TRY(h, h->ids_stage.ensure(size_t(n) * 8));
    CUDA_TRY(h, h->rec_stage.ensure(size_t(n) * D * 4));
    CUDA_TRY(h, cudaMemcpyAsync(h->ids_stage.p, ids, size_t(n) * 8, cudaMemcpyHostToDevice, h->stream));
    d_ids = h->ids_stage.as<long long>();
    d_out = h->rec_stage.as<float>();
  }
  const int warps = 8;
  dim3 grid(unsigned((n + warps - 1) / warps)), block(32 * warps);
  const float* master = h->has_master() ? h->master : nullptr;
  if (h->f16()) gather_rows_kernel<__half><<<grid, block, 0, h->stream>>>(d_ids, n, h->n, h->d, h->dp, master, (const __half*)h->hi, h->id_offset, 0, d_out);
  else gather_rows_kernel<__nv_bfloat16><<<grid, block, 0, h->stream>>>(d_ids, n, h->n, h->d, h->dp, master, (const __nv_bfloat16*)h->hi, h->id_offset, 0, d_out);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if (mem == RDB_MEM_HOST) {
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, size_t(n) * D * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, host_sync(h));
  }
  return RDB_OK;
}

int rdb_reconstruct(rdb_handle* h, int64_t id, float* out) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  if (id - h->id_offset < 0 || id - h->id_offset >= h->n)
    return fail(h, RDB_ERR_INVALID, "reconstruct: id " + std::to_string(id) + " out of range");
  return rdb_reconstruct_batch(h, &id, 1, RDB_MEM_HOST, out);
}

int rdb_set_labels(rdb_handle* h, const float* labels, int64_t n) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (n < 0 || (n > 0 && !labels)) return fail(h, RDB_ERR_INVALID, "set_labels: bad arguments");
  cudaStreamSynchronize(h->stream);
  cudaFree(h->labels); h->labels = nullptr; h->nlabels = 0;
  if (n == 0) return RDB_OK;
  CUDA_TRY(h, cudaMalloc(&h->labels, size_t(n) * 4));
  CUDA_TRY(h, cudaMemcpyAsync(h->labels, labels, size_t(n) * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, host_sync(h));
  h->nlabels = n;
  return RDB_OK;
}

int rdb_filter_first_k(rdb_handle* h, const int64_t* idx, const float* dist, const float* labels, int64_t nq, int ks,
                       const int64_t* row_code, int64_t ntotal, const int64_t* excl_sorted, int n_excl, int K,
                       int64_t* out_idx, float* out_dist, float* out_labels) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (nq < 0 || ks < 0 || K < 1 || ntotal < 0 || !out_idx || !out_dist || !out_labels || (ks > 0 && (!idx || !dist)) ||
      (n_excl > 0 && (!row_code || !excl_sorted)))
    return fail(h, RDB_ERR_INVALID, "filter_first_k: bad arguments");
  if (nq == 0) return RDB_OK;
  filter_first_k_kernel<<<unsigned((nq + 3) / 4), 128, 0, h->stream>>>(      // one warp per query
      reinterpret_cast<const long long*>(idx), dist, labels, int(nq), ks, reinterpret_cast<const long long*>(row_code),
      (long long)ntotal, reinterpret_cast<const long long*>(excl_sorted), n_excl, K,
      reinterpret_cast<long long*>(out_idx), out_dist, out_labels);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

int rdb_label_vote(rdb_handle* h, const float* lbl, int64_t nq, int k, int kvote, int mem, float* vote) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (nq < 0 || k < 1 || !lbl || !vote) return fail(h, RDB_ERR_INVALID, "label_vote: bad arguments");
  if (nq == 0) return RDB_OK;
  const float* d_l = lbl; float* d_v = vote;
  if (mem == RDB_MEM_HOST) {
    CUDA_TRY(h, h->o_lbl.ensure(size_t(nq) * k * 4));
    CUDA_TRY(h, h->o_dist.ensure(size_t(nq) * 4));
    CUDA_TRY(h, cudaMemcpyAsync(h->o_lbl.p, lbl, size_t(nq) * k * 4, cudaMemcpyHostToDevice, h->stream));
    d_l = h->o_lbl.as<float>(); d_v = h->o_dist.as<float>();
  }
  label_vote_kernel<<<unsigned((nq + 255) / 256), 256, 0, h->stream>>>(d_l, int(nq), k, kvote, d_v);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if (mem == RDB_MEM_HOST) {
    CUDA_TRY(h, cudaMemcpyAsync(vote, d_v, size_t(nq) * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, host_sync(h));
  }
  return RDB_OK;
}

int64_t rdb_ntotal(rdb_handle* h) { return h ? h->n : 0; }
int rdb_dim(rdb_handle* h) { return h ? h->d : 0; }
int rdb_metric(rdb_handle* h) { return h ? h->metric : -1; }
int rdb_store_dtype(rdb_handle* h) { return h ? h->store : -1; }
int64_t rdb_launch_count(rdb_handle* h) { return h ? h->launches : 0; }
int64_t rdb_host_sync_count(rdb_handle* h) { return h ? h->host_syncs : 0; }

int rdb_set_id_offset(rdb_handle* h, int64_t offset) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  h->id_offset = offset;
  return RDB_OK;
}

int rdb_last_kernel_ms(rdb_handle* h, float* ms, int* algo, int* nsplits) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  if (!h->ev_valid) return fail(h, RDB_ERR_INVALID, "no search has run yet");
  DeviceGuard dg(h->device);
  CUDA_TRY(h, cudaEventSynchronize(h->ev1));
  float t = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&t, h->ev0, h->ev1));
  if (ms) *ms = t;
  if (algo) *algo = h->last_algo;
  if (nsplits) *nsplits = h->last_S;
  return RDB_OK;
}

int64_t rdb_last_uncertified(rdb_handle* h) {
  if (!h) return 0;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  counts_resolve(h, true);             // the counters travel behind the search: wait for them
  return h->last_uncertified;
}

int rdb_last_tier1(rdb_handle* h, int64_t* queries, int64_t* uncertified, int* candidates) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    DeviceGuard dg(h->device);
    counts_resolve(h, true);
  }
  if (queries) *queries = h->last_tier1_queries;
  if (uncertified) *uncertified = h->last_tier1_uncertified;
  if (candidates) *candidates = h->last_tier1_queries > 0 ? h->last_tier1_kc : 0;
  return RDB_OK;
}

int rdb_mem_info(rdb_handle* h, size_t* index_bytes, size_t* scratch_bytes, size_t* free_bytes, size_t* total_bytes) {
  size_t fr = 0, tot = 0, ib = 0, sb = 0;
  if (h) {
    std::lock_guard<std::mutex> lock(h->mu);
    DeviceGuard dg(h->device);
    CUDA_TRY(h, cudaMemGetInfo(&fr, &tot));
    if (h->has_master()) ib += size_t(h->cap) * h->d * 4;
    ib += size_t(h->cap) * h->dp * 2 * (h->has_lo() ? 2 : 1);
    ib += size_t(h->cap) * 4 + size_t(h->cap / 32) * 4 + size_t(h->nlabels) * 4 + (h->yext ? size_t(h->cap) * 16 : 0);
    for (auto m : kScratch) sb += (h->*m).bytes;
    sb += h->pin_bytes;
  } else {
    CUDA_TRY(nullptr, cudaMemGetInfo(&fr, &tot));
  }
  if (index_bytes) *index_bytes = ib;
  if (scratch_bytes) *scratch_bytes = sb;
  if (free_bytes) *free_bytes = fr;
  if (total_bytes) *total_bytes = tot;
  return RDB_OK;
}

int rdb_release_scratch(rdb_handle* h) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  CUDA_TRY(h, host_sync(h));
  for (auto m : kScratch) (h->*m).release();
  if (h->pin) cudaFreeHost(h->pin);
  h->pin = nullptr; h->pin_bytes = 0;
  h->qext_rows = 0;
  h->ev_valid = false;
  cudaGetLastError();
  return RDB_OK;
}

int rdb_truncate(rdb_handle* h, int64_t n_keep) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (n_keep < 0 || n_keep > h->n) return fail(h, RDB_ERR_INVALID, "truncate: n_keep must be in [0, ntotal]");
  if (n_keep == h->n) return RDB_OK;
  CUDA_TRY(h, host_sync(h));
  // forgotten rows must look like never-filled ones: |y|^2 = 0 keeps the group minima a valid (loose) lower bound
  CUDA_TRY(h, cudaMemsetAsync(h->ynorm + n_keep, 0, size_t(h->n - n_keep) * 4, h->stream));
  const int64_t g0 = n_keep / 32, g1 = (h->n - 1) / 32;
  ynorm_min32_kernel<<<unsigned((g1 - g0 + 1 + 7) / 8), 256, 0, h->stream>>>(h->ynorm, g0, g1 - g0 + 1, h->ynmin32);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, host_sync(h));
  h->n = n_keep;                       // labels (if any) no longer match ntotal -> ignored until set again
  return RDB_OK;
}

int rdb_set_option(rdb_handle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(h, RDB_ERR_INVALID, "set_option: bad arguments");
  std::lock_guard<std::mutex> lock(h->mu);
  const std::string n(name);
  rdb_options& o = h->opt;
  if (n == "tc_cta_group") o.tc_cta_group = int(value);
  else if (n == "tc_lockstep") o.tc_lockstep = int(value);
  else if (n == "tc_lockstep_spins") o.tc_lockstep_spins = int(value);
  else if (n == "tc_stages") o.tc_stages = int(value);
  else if (n == "tc_query_stationary") o.tc_query_stationary = int(value);
  else if (n == "tc_pivot") o.tc_pivot = int(value);
  else if (n == "tc_chunks") o.tc_chunks = int(std::min<int64_t>(std::max<int64_t>(value, 0), 256 / TC_LISTS));
  else if (n == "tier1") o.tier1 = int(value);
  else if (n == "tier1_kc") o.tier1_kc = int(value);
  else if (n == "largek_scorer") o.largek_scorer = int(value);
  else if (n == "largek_rows") o.largek_rows = value;
  else if (n == "largek_sample") o.largek_sample = int(value);
  else if (n == "largek_split") o.largek_split = int(value);
  else if (n == "host_pipeline") o.host_pipeline = int(value);
  else if (n == "tier1_share2") o.tier1_share2 = int(value);
  else if (n == "tc_list10") o.tc_list10 = int(value);
#ifdef RDB_PROFILING
  else if (n == "tc_debug") o.tc_debug = int(value);
  else if (n == "stream_prof") o.stream_prof = value;
#endif
  else return fail(h, RDB_ERR_INVALID, "set_option: unknown option '" + n + "'");
  return RDB_OK;
}

// ---- faiss IndexFlat on-disk layout (faiss/impl/index_write.cpp, v1.10): u32 fourcc, i32 d, i64 ntotal,
//      i64 dummy, i64 dummy, u8 is_trained, i32 metric_type (0 = IP, 1 = L2), u64 count (= ntotal * d floats), data.
static uint32_t fourcc(const char* s) { return uint32_t(uint8_t(s[0])) | uint32_t(uint8_t(s[1])) << 8 | uint32_t(uint8_t(s[2])) << 16 | uint32_t(uint8_t(s[3])) << 24; }

int rdb_copy_async(rdb_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  if (bytes == 0) return RDB_OK;
  if (!dst || !src) return fail(h, RDB_ERR_INVALID, "copy_async: null buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  // A copy KERNEL on this device (peer memory is read with plain loads over NVLink), not cudaMemcpyAsync: a peer DMA
  // copy is ordered by the driver against the SOURCE device's legacy default stream as well, so a pull from a GPU
  // whose worker had already enqueued its 90 ms search waited for that search (8 GPUs: steps of 99 or 150-180 ms
  // depending on which host thread got there first).
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15) == 0;
  const size_t units = vec ? bytes / 16 : bytes;
  const int blocks = int(std::min<size_t>((units + 255) / 256, size_t(h->num_sms) * 8));
  if (vec) copy_bytes_kernel<uint4><<<blocks, 256, 0, h->stream>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), units);
  else copy_bytes_kernel<unsigned char><<<blocks, 256, 0, h->stream>>>(static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), units);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

int rdb_enable_peer_access(rdb_handle* h, int peer_device) {
  if (!h) return fail(nullptr, RDB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  if (peer_device == h->device) return RDB_OK;
  DeviceGuard dg(h->device);
  int can = 0;
  CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, h->device, peer_device));
  if (!can) return fail(h, RDB_ERR_UNSUPPORTED, "device " + std::to_string(h->device) + " cannot access device " +
                                                    std::to_string(peer_device) + " (no NVLink/PCIe peer path)");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  CUDA_TRY(h, e);
  return RDB_OK;
}

int rdb_serialize(rdb_handle* h, const char* path) {
  if (!h || !path) return fail(h, RDB_ERR_INVALID, "serialize: bad arguments");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  FILE* f = fopen(path, "wb");
  if (!f) return fail(h, RDB_ERR_IO, std::string("serialize: cannot open ") + path);
  const uint32_t cc = fourcc(h->metric == RDB_METRIC_IP ? "IxFI" : "IxF2");
  const int32_t d = h->d; const int64_t n = h->n, dummy = 1 << 20; const uint8_t trained = 1;
  const int32_t metric = h->metric == RDB_METRIC_IP ? 0 : 1; const uint64_t count = uint64_t(n) * uint64_t(d);
  bool ok = fwrite(&cc, 4, 1, f) == 1 && fwrite(&d, 4, 1, f) == 1 && fwrite(&n, 8, 1, f) == 1 &&
            fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&trained, 1, 1, f) == 1 &&
            fwrite(&metric, 4, 1, f) == 1 && fwrite(&count, 8, 1, f) == 1;
  const int64_t chunk = std::max<int64_t>(1, (int64_t(64) << 20) / (int64_t(d) * 4));
  std::vector<float> host(size_t(std::min(chunk, std::max<int64_t>(n, 1))) * d);
  for (int64_t r0 = 0; ok && r0 < n; r0 += chunk) {
    const int64_t m = std::min(chunk, n - r0);
    const float* src;
    if (h->has_master()) src = h->master + r0 * d;
    else {
      if (h->rec_stage.ensure(size_t(m) * d * 4) != cudaSuccess) { ok = false; break; }
      const int warps = 8;
      dim3 grid(unsigned((m + warps - 1) / warps)), block(32 * warps);
      if (h->f16()) gather_rows_kernel<__half><<<grid, block, 0, h->stream>>>(nullptr, m, h->n, h->d, h->dp, nullptr, (const __half*)h->hi, 0, r0, h->rec_stage.as<float>());
      else gather_rows_kernel<__nv_bfloat16><<<grid, block, 0, h->stream>>>(nullptr, m, h->n, h->d, h->dp, nullptr, (const __nv_bfloat16*)h->hi, 0, r0, h->rec_stage.as<float>());
      h->launches++;
      src = h->rec_stage.as<float>();
    }
    if (cudaMemcpyAsync(host.data(), src, size_t(m) * d * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
    ok = fwrite(host.data(), 4, size_t(m) * d, f) == size_t(m) * d;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(h, RDB_ERR_IO, std::string("serialize: write to ") + path + " failed");
  return RDB_OK;
}

int rdb_deserialize(const char* path, int store_dtype, int device, unsigned flags, rdb_handle** out) {
  if (!path || !out) return fail(nullptr, RDB_ERR_INVALID, "deserialize: bad arguments");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(nullptr, RDB_ERR_IO, std::string("deserialize: cannot open ") + path);
  uint32_t cc = 0; int32_t d = 0, metric = 0; int64_t n = 0, dummy = 0; uint8_t trained = 0; uint64_t count = 0;
  bool ok = fread(&cc, 4, 1, f) == 1 && fread(&d, 4, 1, f) == 1 && fread(&n, 8, 1, f) == 1 &&
            fread(&dummy, 8, 1, f) == 1 && fread(&dummy, 8, 1, f) == 1 && fread(&trained, 1, 1, f) == 1 &&
            fread(&metric, 4, 1, f) == 1;
  if (ok && metric > 1) { float arg; ok = fread(&arg, 4, 1, f) == 1; }
  ok = ok && fread(&count, 8, 1, f) == 1;
  if (!ok || (cc != fourcc("IxF2") && cc != fourcc("IxFI") && cc != fourcc("IxFl")) || d < 1 || n < 0 ||
      count != uint64_t(n) * uint64_t(d)) {
    fclose(f);
    return fail(nullptr, RDB_ERR_IO, std::string("deserialize: ") + path + " is not a faiss IndexFlat file");
  }
  if (metric != 0 && metric != 1) { fclose(f); return fail(nullptr, RDB_ERR_UNSUPPORTED, "deserialize: metric type not L2/IP"); }
  rdb_handle* h = nullptr;
  int rc = rdb_create(d, metric == 0 ? RDB_METRIC_IP : RDB_METRIC_L2, store_dtype, device, flags, &h);
  if (rc) { fclose(f); return rc; }
  rc = rdb_reserve(h, n);
  const int64_t chunk = std::max<int64_t>(1, (int64_t(64) << 20) / (int64_t(d) * 4));
  std::vector<float> host(size_t(std::min(chunk, std::max<int64_t>(n, 1))) * d);
  for (int64_t r0 = 0; rc == RDB_OK && r0 < n; r0 += chunk) {
    const int64_t m = std::min(chunk, n - r0);
    if (fread(host.data(), 4, size_t(m) * d, f) != size_t(m) * d) { rc = fail(nullptr, RDB_ERR_IO, "deserialize: truncated file"); break; }
    rc = rdb_add(h, host.data(), m, RDB_MEM_HOST, 0);
    if (rc) g_err = h->err;
  }
  fclose(f);
  if (rc) { rdb_destroy(h); return rc; }
  *out = h;
  return RDB_OK;
}

}  // extern "C"
