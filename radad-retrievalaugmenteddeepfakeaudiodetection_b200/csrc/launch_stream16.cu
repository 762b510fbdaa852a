// Small-batch streaming scorer launchers for the 16-bit stores; see launch_stream.cu.
#include "launch_stream_impl.h"

namespace rdb {

int launch_stream16(rdb_handle* h, const StreamParams& p, int blocks, int mode) {
  const bool l2 = h->metric == RDB_METRIC_L2;
  if (h->f16())
    return l2 ? launch_stream_mode<__half, true>(h, p, blocks, mode) : launch_stream_mode<__half, false>(h, p, blocks, mode);
  return l2 ? launch_stream_mode<__nv_bfloat16, true>(h, p, blocks, mode)
            : launch_stream_mode<__nv_bfloat16, false>(h, p, blocks, mode);
}

}  // namespace rdb
