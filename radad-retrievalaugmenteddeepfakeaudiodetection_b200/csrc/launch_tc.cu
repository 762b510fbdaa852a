// tcgen05 scorer launchers (kernel 2, score_tc.cuh): tensor-map encoding + selector / metric / CTA-group dispatch.
#include "handle.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>

#include "launch_tc_impl.h"

namespace rdb {

namespace {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    else
      cudaGetLastError();
  });
  return fn;
}


}  // namespace

int encode_2d(rdb_handle* h, CUtensorMap* m, const void* base, int64_t rows, int D, int Dp, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(h, RDB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cuuint64_t(D), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(Dp) * 2};
  cuuint32_t box[2] = {cuuint32_t(TC_BK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, h->f16() ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RDB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(int(r)));
  return RDB_OK;
}


int launch_tc_cg2(rdb_handle* h, TcParams& p, int k);   // launch_tc2.cu

int launch_tc_cg(rdb_handle* h, TcParams& p, int k, int cg) {
  return cg == 2 ? launch_tc_cg2(h, p, k) : launch_tc_cg_t<1>(h, p, k);
}

}  // namespace rdb
