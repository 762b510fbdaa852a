// Exact CUDA-core scorer launchers (kernel 4, score_simt.cuh): selecting form (the k > 128 DUMP form: launch_simt_dump.cu).
#include "handle.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "score_simt.cuh"

namespace rdb {

namespace {
template <int KT, bool L2, typename T, bool ALIGNED>
int launch_simt_t(rdb_handle* h, const T* Q, const T* Y, int nq, int ld, int nqt, int S, int rows_per_chunk,
                  float* ck, int* ci, int kout, const DevPlan* plan) {
  auto kern = score_select_simt_kernel<KT, L2, T, ALIGNED>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)simt_smem_bytes()));
  // device-sized launch: a persistent grid walks the units the plan describes
  const unsigned grid = plan ? unsigned(3 * h->num_sms) : unsigned(nqt) * unsigned(S);
  kern<<<dim3(grid), dim3(256), simt_smem_bytes(), h->stream>>>(
      Q, Y, h->ynorm, nq, int(h->n), h->d, ld, nqt, S, rows_per_chunk, ck, ci, kout, nullptr, 0ll, 0, plan);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}

template <int KT, bool L2>
int launch_simt_k(rdb_handle* h, const float* qf, const void* qhi, int nq, int nqt, int S, int rows_per_chunk,
                  float* ck, int* ci, int kout, const DevPlan* plan) {
  if (h->store == RDB_STORE_F32) {
    const float* Q = qf;
    if (h->d % 4 == 0) return launch_simt_t<KT, L2, float, true>(h, Q, h->master, nq, h->d, nqt, S, rows_per_chunk, ck, ci, kout, plan);
    return launch_simt_t<KT, L2, float, false>(h, Q, h->master, nq, h->d, nqt, S, rows_per_chunk, ck, ci, kout, plan);
  }
  if (h->f16())
    return launch_simt_t<KT, L2, __half, true>(h, (const __half*)qhi, (const __half*)h->hi, nq, h->dp, nqt, S, rows_per_chunk, ck, ci, kout, plan);
  return launch_simt_t<KT, L2, __nv_bfloat16, true>(h, (const __nv_bfloat16*)qhi, (const __nv_bfloat16*)h->hi, nq, h->dp, nqt, S, rows_per_chunk, ck, ci, kout, plan);
}

}  // namespace

int launch_simt(rdb_handle* h, const float* qf, const void* qhi, int nq, int k, int nqt, int S, int rows_per_chunk,
                float* ck, int* ci, const DevPlan* plan) {
  if (h->store != RDB_STORE_F32 && h->cur_hscale != 1.0f)
    return fail(h, RDB_ERR_INVALID, "internal: the CUDA-core scorer was handed queries staged for the tensor-core norm slice");
  const bool l2 = h->metric == RDB_METRIC_L2;
#define SIMT_CASE(KT)                                                                          \
  return l2 ? launch_simt_k<KT, true>(h, qf, qhi, nq, nqt, S, rows_per_chunk, ck, ci, k, plan) \
            : launch_simt_k<KT, false>(h, qf, qhi, nq, nqt, S, rows_per_chunk, ck, ci, k, plan)
  if (k <= 16) { SIMT_CASE(16); }
  if (k <= 32) { SIMT_CASE(32); }
  if (k <= 64) { SIMT_CASE(64); }
  SIMT_CASE(128);
#undef SIMT_CASE
}


}  // namespace rdb
