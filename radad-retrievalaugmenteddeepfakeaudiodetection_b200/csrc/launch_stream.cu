// Small-batch streaming scorer launchers (kernel 3, score_stream.cuh).
#include "handle.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "launch_stream_impl.h"

namespace rdb {

int launch_stream16(rdb_handle* h, const StreamParams& p, int blocks, int mode);   // launch_stream16.cu

int launch_stream(rdb_handle* h, StreamParams& p, int blocks, int mode) {
  const bool l2 = h->metric == RDB_METRIC_L2;
  p.metric_l2 = l2 ? 1 : 0;
  int nvec;
  if (h->store == RDB_STORE_F32) { p.Y = h->master; p.ld = h->d; nvec = p.ld / 4; }
  else { p.Y = h->hi; p.ld = h->dp; nvec = p.ld / 8; }
  // lanes per row: whole warp for long rows, 16 / 8 lanes when a row is only a few 128-bit vectors
  p.lpr_log2 = nvec >= 96 ? 5 : (nvec >= 48 ? 4 : 3);
  if (h->store == RDB_STORE_F32)
    return l2 ? launch_stream_mode<float, true>(h, p, blocks, mode) : launch_stream_mode<float, false>(h, p, blocks, mode);
  return launch_stream16(h, p, blocks, mode);
}


}  // namespace rdb
