// Exact CUDA-core scorer, DUMP form (k > 128: dense fp32 keys for select_dense_kernel); see launch_simt.cu.
#include "handle.h"

#include "score_simt.cuh"

namespace rdb {

namespace {
template <bool L2, typename T, bool ALIGNED>
int launch_simt_dump_t(rdb_handle* h, const T* Q, const T* Y, int nq, int ld, int nqt, int S, int rows_per_chunk,
                       int row0, int row_end, float* dump, long long pitch) {
  auto kern = score_select_simt_kernel<16, L2, T, ALIGNED, true>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)simt_smem_bytes()));
  kern<<<dim3(unsigned(nqt) * unsigned(S)), dim3(256), simt_smem_bytes(), h->stream>>>(
      Q, Y, h->ynorm, nq, row_end, h->d, ld, nqt, S, rows_per_chunk, nullptr, nullptr, 0, dump, pitch, row0, nullptr);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

template <bool L2>
int launch_simt_dump_l2(rdb_handle* h, const QueryView& qv, int q0, int nq, int nqt, int S, int rows_per_chunk, int row0,
                     int row_end, float* dump, long long pitch) {
  if (h->store == RDB_STORE_F32) {
    const float* Q = qv.qf + size_t(q0) * h->d;
    if (h->d % 4 == 0) return launch_simt_dump_t<L2, float, true>(h, Q, h->master, nq, h->d, nqt, S, rows_per_chunk, row0, row_end, dump, pitch);
    return launch_simt_dump_t<L2, float, false>(h, Q, h->master, nq, h->d, nqt, S, rows_per_chunk, row0, row_end, dump, pitch);
  }
  if (h->f16())
    return launch_simt_dump_t<L2, __half, true>(h, (const __half*)qv.qhi + size_t(q0) * h->dp, (const __half*)h->hi, nq, h->dp, nqt, S, rows_per_chunk, row0, row_end, dump, pitch);
  return launch_simt_dump_t<L2, __nv_bfloat16, true>(h, (const __nv_bfloat16*)qv.qhi + size_t(q0) * h->dp, (const __nv_bfloat16*)h->hi, nq, h->dp, nqt, S, rows_per_chunk, row0, row_end, dump, pitch);
}

}  // namespace

int launch_simt_dump(rdb_handle* h, const QueryView& qv, int q0, int nq, int nqt, int S, int rows_per_chunk, int row0,
                     int row_end, float* dump, long long pitch) {
  if (h->store != RDB_STORE_F32 && h->cur_hscale != 1.0f)
    return fail(h, RDB_ERR_INVALID, "internal: the CUDA-core scorer was handed queries staged for the tensor-core norm slice");
  return h->metric == RDB_METRIC_L2
             ? launch_simt_dump_l2<true>(h, qv, q0, nq, nqt, S, rows_per_chunk, row0, row_end, dump, pitch)
             : launch_simt_dump_l2<false>(h, qv, q0, nq, nqt, S, rows_per_chunk, row0, row_end, dump, pitch);
}


}  // namespace rdb
