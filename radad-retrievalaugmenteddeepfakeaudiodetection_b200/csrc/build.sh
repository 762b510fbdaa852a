#!/usr/bin/env bash
# Build libradad_flat.so (C ABI of include/radad_flat.h) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libradad_flat.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function -shared \
  ${RDB_PTXAS_V:+-Xptxas -v} \
  -o "${OUT}" "${HERE}/radad_flat.cu"
echo "built ${OUT}"
