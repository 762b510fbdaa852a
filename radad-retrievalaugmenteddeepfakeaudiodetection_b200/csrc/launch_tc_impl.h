// Template part of the tcgen05 launchers, compiled once per CTA-group size (launch_tc.cu: cta_group::1,
// launch_tc2.cu: cta_group::2) so the two halves of the instantiations build in parallel.
#pragma once
#include "handle.h"

#include <algorithm>
#include <mutex>

#include "score_tc.cuh"

namespace rdb {
namespace {

constexpr int kReservoirCap = 320;   // large-k epilogue: room for 128 kept + >= 160 appended between prunes

template <class SEL, bool L2V, int CG>
int launch_tc_kernel(rdb_handle* h, const TcParams& p, int groups) {
  auto kern = score_select_tc_kernel<SEL, L2V, CG>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<CG>()));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(groups * CG));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = tc_smem_bytes<CG>();
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(h, cudaLaunchKernelEx(&cfg, kern, p));
  return RDB_OK;
}

template <int CG>
int launch_tc_cg_t(rdb_handle* h, TcParams& p, int k) {
  int rc;
  if ((rc = encode_2d(h, &p.tmap_y[0], h->hi, h->n, h->d, h->dp, TC_BN / CG))) return rc;
  if (p.nterms == 3) { if ((rc = encode_2d(h, &p.tmap_y[1], h->lo, h->n, h->d, h->dp, TC_BN / CG))) return rc; }
  else p.tmap_y[1] = p.tmap_y[0];
  if (p.ext && (rc = encode_2d(h, &p.tmap_yx, h->yext, h->n, 8, 8, TC_BN / CG))) return rc;
  p.idesc = make_idesc_f16(TC_BM * CG, TC_BN, h->f16() ? 0 : 1);
  const int groups = std::min(p.num_units, h->num_sms / CG);
  const bool l2 = h->metric == RDB_METRIC_L2 && !p.ext;     // norm slice: the accumulator already is the L2 key
  if (p.dump)       return l2 ? launch_tc_kernel<SelectDump, true, CG>(h, p, groups) : launch_tc_kernel<SelectDump, false, CG>(h, p, groups);
  if (p.share2 && k <= 16)
    return l2 ? launch_tc_kernel<SelectSmall<16, 2>, true, CG>(h, p, groups) : launch_tc_kernel<SelectSmall<16, 2>, false, CG>(h, p, groups);
  // k <= 10 (BASELINE's k): a 10-entry list -- the admission threshold is the 10th key instead of the 16th (a third fewer
  // insertions) and an insertion is 9 compare-exchange steps instead of 15; option "tc_list10" = 0 keeps 16 entries
  if (k <= 10 && h->opt.tc_list10)
    return l2 ? launch_tc_kernel<SelectSmall<10>, true, CG>(h, p, groups) : launch_tc_kernel<SelectSmall<10>, false, CG>(h, p, groups);
  if (k <= 16)      return l2 ? launch_tc_kernel<SelectSmall<16>, true, CG>(h, p, groups) : launch_tc_kernel<SelectSmall<16>, false, CG>(h, p, groups);
  else if (k <= 32) return l2 ? launch_tc_kernel<SelectSmall<32>, true, CG>(h, p, groups) : launch_tc_kernel<SelectSmall<32>, false, CG>(h, p, groups);
  return l2 ? launch_tc_kernel<SelectReservoir<kReservoirCap>, true, CG>(h, p, groups)
            : launch_tc_kernel<SelectReservoir<kReservoirCap>, false, CG>(h, p, groups);
}

}  // namespace
}  // namespace rdb
