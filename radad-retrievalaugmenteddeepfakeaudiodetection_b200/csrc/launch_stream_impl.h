// Template part of the streaming-scorer launchers, compiled once per store type family (launch_stream.cu: fp32,
// launch_stream16.cu: bf16 / f16).
#pragma once
#include "handle.h"

#include "score_stream.cuh"

namespace rdb {
namespace {
template <typename T, int NQ, bool L2, int MODE>
int launch_stream_kernel(rdb_handle* h, const StreamParams& p, int blocks, size_t smem) {
  auto kern = score_select_stream_kernel<T, NQ, L2, MODE>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(blocks), dim3(STREAM_THREADS), smem, h->stream>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return RDB_OK;
}
template <typename T, bool L2>
int launch_stream_mode(rdb_handle* h, const StreamParams& p, int blocks, int mode) {
  const int nqt = p.nq <= 1 ? 1 : (p.nq <= 2 ? 2 : 4);
  const size_t smem = stream_smem_bytes(nqt, p.ld, mode, p.np.nleaves, p.q_raw ? 0 : p.nq * p.D);
#define STREAM_NQ(MODE)                                                                         \
  (nqt == 1 ? launch_stream_kernel<T, 1, L2, MODE>(h, p, blocks, smem)                          \
            : (nqt == 2 ? launch_stream_kernel<T, 2, L2, MODE>(h, p, blocks, smem)              \
                        : launch_stream_kernel<T, 4, L2, MODE>(h, p, blocks, smem)))
  if (mode == STREAM_LIST1) return STREAM_NQ(STREAM_LIST1);
  if (mode == STREAM_LIST4) return STREAM_NQ(STREAM_LIST4);
  return STREAM_NQ(STREAM_FILTER);
#undef STREAM_NQ
}
}  // namespace
}  // namespace rdb
