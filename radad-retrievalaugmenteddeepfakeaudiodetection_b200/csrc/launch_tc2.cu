// tcgen05 scorer launchers, CTA-pair form (cta_group::2); see launch_tc.cu.
#include "launch_tc_impl.h"

namespace rdb {

int launch_tc_cg2(rdb_handle* h, TcParams& p, int k) { return launch_tc_cg_t<2>(h, p, k); }

}  // namespace rdb
