/* radad_flat.h -- C ABI of the B200-native exact flat nearest-neighbour index.
 *
 * This is the drop-in boundary for the RADAD retrieval hot path.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository; FAISS is the reference's third-party
 * dependency, pinned `faiss-gpu-cu11==1.10.0`, requirements.txt:11).  The reference binds that dependency
 * through SWIG; a maintainer binds this library through ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross the boundary;
 *   - every function returns 0 on success, non-zero on failure (RDB_ERR_*); the message is available from
 *     rdb_last_error(handle) (or rdb_last_error(NULL) for failures that produced no handle);
 *   - `mem` says where caller-owned buffers live: RDB_MEM_HOST (the reference's numpy path: copies are done
 *     inside the call and it returns when the results are in the caller's buffers) or RDB_MEM_DEVICE
 *     (CUDA device pointers on the handle's device; work is enqueued on the handle's stream, which the
 *     caller synchronises -- rdb_set_stream / rdb_sync);
 *   - vectors are row-major float32 [n, d], C-contiguous -- exactly what vector_database.py:118-119,166-167
 *     guarantees before it calls faiss;
 *   - ids are 0-based insertion order, int64, plus the handle's id offset (multi-GPU row shards);
 *   - results are sorted best-first: squared L2 ascending / inner product descending (faiss IndexFlat
 *     semantics); ties on distance are broken by the lowest id;
 *   - there is NO CPU fallback: without a CUDA device rdb_create fails.
 */
#ifndef RADAD_FLAT_H
#define RADAD_FLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdb_handle rdb_handle;

enum { RDB_METRIC_L2 = 0, RDB_METRIC_IP = 1 };                       /* faiss.IndexFlatL2 / IndexFlatIP   */
enum { RDB_STORE_F32 = 0, RDB_STORE_BF16 = 1, RDB_STORE_F16 = 2 };   /* GpuIndexFlatConfig.useFloat16 -> 2 */
enum { RDB_MEM_HOST = 0, RDB_MEM_DEVICE = 1 };
enum { RDB_ALGO_AUTO = 0, RDB_ALGO_SIMT = 1, RDB_ALGO_TC = 2, RDB_ALGO_STREAM = 3 }; /* scorer (tests / bench) */
enum {
  RDB_OK = 0, RDB_ERR_INVALID = 1, RDB_ERR_CUDA = 2, RDB_ERR_NOMEM = 3, RDB_ERR_IO = 4, RDB_ERR_UNSUPPORTED = 5
};
/* rdb_create flags */
enum { RDB_FLAG_KEEP_F32_MASTER = 1 /* 16-bit stores also keep the fp32 rows: exact reconstruct */ };

/* ABI version of this header; bumped on any signature change. */
int rdb_abi_version(void);

/* Replaces faiss.StandardGpuResources() + faiss.GpuIndexFlatL2/IP(res, d, cfg) -- vector_database.py:39-45,
 * 62-64,78-84 (create_index).  `device` < 0 means the current CUDA device. */
int rdb_create(int d, int metric, int store_dtype, int device, unsigned flags, rdb_handle** out);

/* Replaces the index destructor / cleanup_gpu_resources -- vector_database.py:259-273. */
int rdb_destroy(rdb_handle* h);

/* Last error message of this handle (or of the calling thread when h == NULL).  Never NULL. */
const char* rdb_last_error(rdb_handle* h);

/* Use a caller-owned CUDA stream (cudaStream_t) for all work of this handle.  The value is used as is: NULL is
 * the legacy default stream (what torch.cuda.current_stream().cuda_stream returns by default).
 * rdb_use_own_stream goes back to the handle's private non-blocking stream (the state after rdb_create).
 * rdb_sync blocks until the handle's current stream is idle. */
int rdb_set_stream(rdb_handle* h, void* cuda_stream);
int rdb_use_own_stream(rdb_handle* h);
int rdb_sync(rdb_handle* h);

/* Pre-size device storage for `n_total` rows (optional; add grows geometrically otherwise). */
int rdb_reserve(rdb_handle* h, int64_t n_total);

/* Replaces index.add(x) -- vector_database.py:138 -- fused with _maybe_normalize (:100-105) when
 * `normalize` != 0: rows are L2-normalised as x / (|x| + 1e-12), converted to the store dtype, and |y|^2 of
 * the stored values is cached, in one pass on the device. */
int rdb_add(rdb_handle* h, const float* x, int64_t n, int mem, int normalize);

/* Replaces index.search(q, k) -- vector_database.py:181 -- fused with the query-side _maybe_normalize
 * (:166).  out_dist float32[nq,k], out_idx int64[nq,k]; out_labels (optional, may be NULL) float32[nq,k]
 * receives labels[id] of every neighbour (the kNN label evidence; 0 when no labels are set).
 * Slots beyond ntotal get id -1 and +inf (L2) / -inf (IP), as faiss does. */
int rdb_search(rdb_handle* h, const float* q, int64_t nq, int k, int mem, int normalize, float* out_dist,
               int64_t* out_idx, float* out_labels);

/* Same, with an explicit scorer (RDB_ALGO_*): SIMT = exact fp32 CUDA-core kernel, TC = tcgen05 kernel (16-bit
 * stores; fp32 stores: split-precision + exact re-rank + certificate), STREAM = small-batch HBM-streaming kernel. */
int rdb_search_algo(rdb_handle* h, const float* q, int64_t nq, int k, int mem, int normalize, int algo,
                    float* out_dist, int64_t* out_idx, float* out_labels);

/* Multi-GPU row shards: per-shard candidates in merge form.  out_key float32[nq,k] (larger is better:
 * IP -> q.y, L2 -> 2 q.y - |y|^2), out_idx int64[nq,k] global ids (-1 = none), out_labels float32[nq,k],
 * out_qnorm float32[nq] (|q|^2 after normalisation).  Device pointers only.  No reference counterpart
 * (the reference is single-GPU: vector_database.py:23). */
int rdb_search_shard(rdb_handle* h, const float* q_dev, int64_t nq, int k, int normalize, float* out_key,
                     int64_t* out_idx, float* out_labels, float* out_qnorm);

/* Final on-device merge of `nlists` per-shard candidate lists laid out [nq][nlists][k] (after the NCCL
 * all-gather): produces what rdb_search would have produced on the unsharded database. */
int rdb_merge_shards(rdb_handle* h, const float* key, const int64_t* idx, const float* labels, int64_t nq,
                     int nlists, int k, const float* qnorm, float* out_dist, int64_t* out_idx,
                     float* out_labels);

/* Fused exchange + merge over NVLink peer memory (the B200-native form of the step above): every rank leaves its
 * candidates in a buffer obtained from rdb_ipc_alloc, peers map it with rdb_ipc_open (CUDA IPC handle, 64 bytes,
 * exchanged once), and ONE kernel merges list g by loading it straight from GPU g's memory -- no all-gather.
 * key_ptrs/idx_ptrs/lbl_ptrs: host arrays of `nlists` device pointers to [nq][k] planes (float32/int64/float32);
 * the caller orders the launch after all ranks finished writing (a stream-ordered barrier).  No reference
 * counterpart. */
int rdb_ipc_alloc(rdb_handle* h, size_t bytes, void** dev_ptr, unsigned char* handle_out /*[64]*/);
int rdb_ipc_open(rdb_handle* h, const unsigned char* handle /*[64]*/, void** dev_ptr);
int rdb_ipc_close(rdb_handle* h, void* dev_ptr);
int rdb_ipc_free(rdb_handle* h, void* dev_ptr);
int rdb_merge_shards_peer(rdb_handle* h, const void* const* key_ptrs, const void* const* idx_ptrs,
                          const void* const* lbl_ptrs, int nlists, int64_t nq, int k, const float* qnorm,
                          float* out_dist, int64_t* out_idx, float* out_labels);

/* Single-process multi-GPU (one host thread drives every shard, as pipeline.py:90 constructs ONE VectorDatabase):
 * let kernels of this handle's device dereference memory of `peer_device` (cudaDeviceEnablePeerAccess; already
 * enabled / same device = success), so rdb_merge_shards_peer can take plain pointers of the other shards' buffers.
 * No reference counterpart. */
int rdb_enable_peer_access(rdb_handle* h, int peer_device);

/* Stream-ordered copy of `bytes` bytes between two device buffers, either of which may live on a peer GPU (after
 * rdb_enable_peer_access), enqueued on THIS handle's stream and device -- a pull by the destination's worker.  The
 * single-process multi-GPU search exchanges the query slices with it: a framework-level cross-device tensor copy is
 * enqueued on the SOURCE device's stream, where it queues up behind that GPU's own 90 ms search (measured on 8 GPUs:
 * two of the eight workers started their search only after a peer had finished its own).  No reference counterpart. */
int rdb_copy_async(rdb_handle* h, void* dst, const void* src, size_t bytes);

/* Replaces index.reconstruct(i) -- pipeline.py:503.  `out` is a HOST float32[d]. */
int rdb_reconstruct(rdb_handle* h, int64_t id, float* out);

/* Batched reconstruct (one kernel instead of B*K faiss calls -- pipeline.py:491-509): out float32[n,d];
 * ids < 0 or out of range give a zero row (the caller's padding value, pipeline.py:511-512). */
int rdb_reconstruct_batch(rdb_handle* h, const int64_t* ids, int64_t n, int mem, float* out);

/* The caller's rank-ordered self-exclusion + "first K survivors of K+10" compaction -- pipeline.py:491-520 -- on the
 * device.  idx/dist/labels [nq][ks] are search results (device), row_code[ntotal] an integer code per row (the file
 * basename the reference compares), excl_sorted[n_excl] the ascending codes to skip.  `ntotal` is the number of rows of
 * the WHOLE database (= the length of row_code): ids outside [0, ntotal) are skipped -- with row shards that is the global
 * row count, not this handle's.  Outputs [nq][K]; missing slots get id -1, label 0, distance NaN (the reference's
 * padding).  All pointers are device pointers. */
int rdb_filter_first_k(rdb_handle* h, const int64_t* idx, const float* dist, const float* labels, int64_t nq, int ks,
                       const int64_t* row_code, int64_t ntotal, const int64_t* excl_sorted, int n_excl, int K,
                       int64_t* out_idx, float* out_dist, float* out_labels);

/* Neighbour labels for the kNN label vote: labels float32[n] (host), n must equal ntotal at search time.
 * Mirrors VectorDatabase.vector_labels -- vector_database.py:16,142. */
int rdb_set_labels(rdb_handle* h, const float* labels, int64_t n);

/* sum over the first kvote neighbours of out_labels[nq,k] -> vote float32[nq] (same `mem` for both). */
int rdb_label_vote(rdb_handle* h, const float* labels_nq_k, int64_t nq, int k, int kvote, int mem, float* vote);

/* index.ntotal / index.d -- vector_database.py:151,169,210; pipeline.py:465,1039; app.py:75,246. */
int64_t rdb_ntotal(rdb_handle* h);
int rdb_dim(rdb_handle* h);
int rdb_metric(rdb_handle* h);
int rdb_store_dtype(rdb_handle* h);

/* Global id of local row 0 (row shards); ids returned = local id + offset. */
int rdb_set_id_offset(rdb_handle* h, int64_t offset);

/* Replaces faiss.index_gpu_to_cpu + faiss.write_index -- vector_database.py:200,203.  Writes the faiss
 * IndexFlat on-disk layout (fourcc IxF2 / IxFI, fp32 rows). */
int rdb_serialize(rdb_handle* h, const char* path);

/* Replaces faiss.read_index + faiss.index_cpu_to_gpu -- vector_database.py:230,233. */
int rdb_deserialize(const char* path, int store_dtype, int device, unsigned flags, rdb_handle** out);

/* Device memory: bytes of row storage owned by this index, bytes of its grow-only search scratch (the counterpart of
 * faiss.StandardGpuResources' temporary memory, vector_database.py:39-45), and free/total of the device
 * (get_gpu_memory_usage -- vector_database.py:245-256).  h == NULL reports 0 / 0 and the current device. */
int rdb_mem_info(rdb_handle* h, size_t* index_bytes, size_t* scratch_bytes, size_t* free_bytes, size_t* total_bytes);

/* Free the search scratch (staging buffers, candidate lists, pinned host staging); the index stays searchable and the
 * scratch regrows on demand.  With rdb_destroy this is what cleanup_gpu_resources (vector_database.py:259-268) maps to. */
int rdb_release_scratch(rdb_handle* h);

/* Forget the rows beyond the first n_keep (n_keep <= ntotal; storage is kept for re-use).  Used to roll a partially
 * applied multi-shard add back so that index.add stays all-or-nothing as in faiss (vector_database.py:138,147-149). */
int rdb_truncate(rdb_handle* h, int64_t n_keep);

/* Per-handle tuning / test options; every default is the production path and the library never reads the process
 * environment.  Names: "tc_cta_group" (0 auto | 1 | 2), "tc_lockstep" (window in groups of 8 tiles, 0 = off),
 * "tc_lockstep_spins", "tc_stages", "tc_chunks" (0 = cost model), "tc_query_stationary" (0 | 1), "tc_pivot" (0 | 1), "tier1" (0 | 1), "tier1_kc"
 * (0 auto | 32 | 64 | 128), "largek_scorer" (0 auto | 1 CUDA cores | 2 tensor cores), "largek_rows" (rows per dense key
 * chunk, 0 = default), "largek_sample" (0 | 1), "largek_split" (0 | 1: split-precision tensor-core keys for fp32 stores),
 * "tc_list10" (0 | 1: 10- or 16-entry register lists for k <= 10),
 * "tier1_share2" (0 | 1: tier 1 keeps its 32 candidates as a two-list cover of 16-entry lists; results identical),
 * "host_pipeline" (0 | 1: searches with HOST buffers of >= 4096 queries / 8 MB upload the batch in pieces on a second
 * stream so that the copy of piece i + 1 overlaps the search of piece i; results are identical either way).
 * Unknown names fail with RDB_ERR_INVALID.  No reference counterpart. */
int rdb_set_option(rdb_handle* h, const char* name, int64_t value);

/* Number of kernels this handle has launched since creation (bench.py's `gpu_launches`). */
int64_t rdb_launch_count(rdb_handle* h);

/* How often a call on this handle blocked the host on its stream (cudaStreamSynchronize).  RDB_MEM_HOST calls block by
 * contract (results land in the caller's buffers); RDB_MEM_DEVICE searches with k <= 104 must not -- every data-dependent
 * step of the certified fp32 search is sized on the device (diagnostic; no reference counterpart). */
int64_t rdb_host_sync_count(rdb_handle* h);

/* Elapsed milliseconds (CUDA events on the handle's stream) of the scoring kernel of the last search, and
 * its name ("tc" / "simt"); 0 on success. */
int rdb_last_kernel_ms(rdb_handle* h, float* ms, int* algo, int* nsplits);

/* fp32 stores are searched by a split-precision tensor-core pass + exact fp32 re-rank; queries whose candidate
 * set cannot be certified against the scorer's error bound are re-searched by the exact CUDA-core kernel.
 * Returns how many queries of the last search took that fallback. */
int64_t rdb_last_uncertified(rdb_handle* h);

/* fp32 stores, k <= 64, >= 262144 rows: a cheaper certified pass runs first (tier 1: ONE tensor-core term on the bf16
 * roundings, 32 or 128 candidates, exact fp32 re-rank, certificate against the bf16 error bound); only the queries it
 * cannot certify go through the three-term pass above.  Returns, for the last search, how many queries entered tier 1,
 * how many of them it could not certify, and the candidates kept per query (0 / 0 / 0 when tier 1 did not run).
 * No reference counterpart (diagnostic). */
int rdb_last_tier1(rdb_handle* h, int64_t* queries, int64_t* uncertified, int* candidates);

#ifdef __cplusplus
}
#endif
#endif /* RADAD_FLAT_H */
